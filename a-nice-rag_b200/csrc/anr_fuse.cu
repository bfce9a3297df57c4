// Weighted reciprocal-rank fusion on device.
//
// Replaces SearchEngine.weighted_reciprocal_rank_fusion (src/search_engine.py:21-34):
//     score[id] += weight[list] * (1 / (k + rank))        rank = 1, 2, ...
// accumulated list after list in float64, then a STABLE sort by score descending
// (ties keep first-insertion order: list 0's entries first, each list in rank order).
// The device result is bit-identical to the Python loop: the reciprocal is an IEEE
// float64 division, the product and the running sum are separate round-to-nearest
// operations (no FMA contraction) and the sum runs in insertion order.
//
// One CTA per query; the working arrays live in shared memory up to 8192 entries per query and
// in an L2-resident global scratch beyond that (the evaluator fuses up to 5 lists x 12 000 ids,
// src/retrieval_eval.py:142-143):
//   1. entries (id, position) with position = insertion index (list-major) are
//      bitonic-sorted ascending by (id, position);
//   2. the first entry of every id run ("head") walks its run in position order and
//      accumulates the float64 score; its position is the id's first insertion;
//   3. (orderable(score), first position) pairs are bitonic-sorted: score descending,
//      position ascending -- exactly the order of Python's stable sorted(reverse=True);
//   4. the first min(top_n, #ids) entries are written out.
#include "anr_internal.h"
#include "anr_common.cuh"

namespace anr {

constexpr int kWrrfThreads = 512;
constexpr int kWrrfMaxEntries = 8192;  // padded (power of two) entries per query in smem

int wrrf_max_entries() { return kWrrfMaxEntries; }

__device__ __forceinline__ uint64_t f64_to_ord(double d) {
  const uint64_t u = static_cast<uint64_t>(__double_as_longlong(d));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_to_f64(uint64_t o) {
  const uint64_t u = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
  return __longlong_as_double(static_cast<long long>(u));
}

// a[] ascending (single 64-bit key)
__device__ __forceinline__ void bitonic_asc_u64(uint64_t* a, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        const uint64_t x = a[lo], y = a[hi];
        if ((x > y) == asc) { a[lo] = y; a[hi] = x; }
      }
    }
  }
  __syncthreads();
}

// pairs (s[], a[]): s descending, ties by low 32 bits of a ascending
__device__ __forceinline__ bool pair_before(uint64_t s0, uint64_t a0, uint64_t s1, uint64_t a1) {
  return s0 > s1 || (s0 == s1 && static_cast<uint32_t>(a0) < static_cast<uint32_t>(a1));
}
__device__ __forceinline__ void bitonic_pairs(uint64_t* s, uint64_t* a, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool fwd = ((lo & size) == 0);
        const uint64_t s0 = s[lo], a0 = a[lo], s1 = s[hi], a1 = a[hi];
        // in a forward block the better pair must sit at lo
        const bool hi_better = pair_before(s1, a1, s0, a0);
        const bool lo_better = pair_before(s0, a0, s1, a1);
        if (fwd ? hi_better : lo_better) { s[lo] = s1; a[lo] = a1; s[hi] = s0; a[hi] = a0; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kWrrfThreads)
wrrf_fuse_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ lens,
                 const double* __restrict__ weights, int n_lists, int list_stride, double rrf_k,
                 int top_n, int n_pow2, uint64_t* __restrict__ scratch,
                 int32_t* __restrict__ out_ids, double* __restrict__ out_scores,
                 int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char smem[];
  // (id << 32) | position, then orderable score of heads (0 otherwise)
  uint64_t* a = scratch ? scratch + static_cast<size_t>(blockIdx.x) * 2 * n_pow2
                        : reinterpret_cast<uint64_t*>(smem);
  uint64_t* s = a + n_pow2;
  __shared__ int list_base[65];                     // insertion offset of each list (n_lists <= 64)
  __shared__ int n_unique;

  const int q = blockIdx.x;
  const int32_t* qids = ids + static_cast<int64_t>(q) * n_lists * list_stride;
  const int32_t* qlens = lens + static_cast<int64_t>(q) * n_lists;
  if (threadIdx.x == 0) {
    int run = 0;
    for (int l = 0; l < n_lists; ++l) {
      list_base[l] = run;
      int len = qlens[l];
      len = len < 0 ? 0 : (len > list_stride ? list_stride : len);
      run += len;
    }
    list_base[n_lists] = run;
    n_unique = 0;
  }
  __syncthreads();
  const int total = list_base[n_lists];

  // 1. entries, padded with the largest key
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) a[i] = ~0ull;
  __syncthreads();
  for (int l = 0; l < n_lists; ++l) {
    const int base = list_base[l], len = list_base[l + 1] - base;
    for (int r = threadIdx.x; r < len; r += blockDim.x)
      a[base + r] = (static_cast<uint64_t>(static_cast<uint32_t>(qids[l * list_stride + r])) << 32) |
                    static_cast<uint32_t>(base + r);
  }
  bitonic_asc_u64(a, n_pow2);

  // 2. heads accumulate their run in insertion order
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    uint64_t sc = 0ull;
    if (i < total) {
      const uint32_t id = static_cast<uint32_t>(a[i] >> 32);
      const bool head = (i == 0) || (static_cast<uint32_t>(a[i - 1] >> 32) != id);
      if (head) {
        double acc = 0.0;
        for (int j = i; j < total && static_cast<uint32_t>(a[j] >> 32) == id; ++j) {
          const int pos = static_cast<int>(static_cast<uint32_t>(a[j]));
          int l = 0;
          while (pos >= list_base[l + 1]) ++l;
          const int rank = pos - list_base[l] + 1;
          const double recip = __ddiv_rn(1.0, __dadd_rn(rrf_k, static_cast<double>(rank)));
          acc = __dadd_rn(acc, __dmul_rn(weights[l], recip));
        }
        sc = f64_to_ord(acc);
        atomicAdd(&n_unique, 1);
      }
    }
    s[i] = sc;
  }
  // 3. order by (score desc, first position asc); non-heads (s == 0) sink to the end
  bitonic_pairs(s, a, n_pow2);

  // 4. emit
  const int n_out = n_unique < top_n ? n_unique : top_n;
  for (int i = threadIdx.x; i < top_n; i += blockDim.x) {
    const int64_t o = static_cast<int64_t>(q) * top_n + i;
    if (i < n_out) {
      out_ids[o] = static_cast<int32_t>(static_cast<uint32_t>(a[i] >> 32));
      out_scores[o] = ord_to_f64(s[i]);
    } else {
      out_ids[o] = -1;
      out_scores[o] = 0.0;
    }
  }
  if (threadIdx.x == 0 && out_counts) out_counts[q] = n_out;
}

// Unions of at most 64 entries (the hybrid query: 2 lists x 10..32 ids): one WARP per query, every
// lane owns up to two entries, all-pairs comparisons through shuffles -- no shared memory, no
// barriers (the block-wide sorts above cost ~10 us of latency even for 20 entries).  Same
// arithmetic: float64 division, product and running sum as separate roundings, summed in
// insertion order; rank = heads with a larger score, or an equal score and an earlier insertion.
constexpr int kWrrfSmallCap = 64;
constexpr int kWrrfSmallWarps = 8;

// One warp, one query: qids [n_lists][list_stride] / qlens [n_lists] may live in global or shared
// memory (the sharded merge below hands over lists it has just built in shared memory).
__device__ __forceinline__ void
wrrf_small_warp(const int32_t* qids, const int32_t* qlens, const double* weights, int n_lists,
                int list_stride, double rrf_k, int top_n, int q, int lane,
                int32_t* __restrict__ out_ids, double* __restrict__ out_scores,
                int32_t* __restrict__ out_counts) {
  // my entries: insertion positions lane and lane + 32
  int id[2] = {0, 0};
  double term[2] = {0.0, 0.0};
  bool valid[2] = {false, false};
  int total = 0;
  for (int l = 0; l < n_lists; ++l) {
    int len = qlens[l];
    len = len < 0 ? 0 : (len > list_stride ? list_stride : len);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int pos = lane + 32 * e;
      if (pos >= total && pos < total + len) {
        const int r = pos - total;
        id[e] = qids[l * list_stride + r];
        term[e] = __dmul_rn(weights[l], __ddiv_rn(1.0, __dadd_rn(rrf_k, static_cast<double>(r + 1))));
        valid[e] = true;
      }
    }
    total += len;
  }
  // scores in insertion order; an entry is a head when no earlier entry carries its id
  double acc[2] = {0.0, 0.0};
  bool head[2] = {valid[0], valid[1]};
  for (int j = 0; j < total; ++j) {
    const int owner = j & 31, slot = j >> 5;
    const int idj = __shfl_sync(kFullMask, slot ? id[1] : id[0], owner);
    const double tj = __shfl_sync(kFullMask, slot ? term[1] : term[0], owner);
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (valid[e] && idj == id[e]) {
        if (j < lane + 32 * e) head[e] = false;
        acc[e] = __dadd_rn(acc[e], tj);
      }
  }
  // rank among the heads: score descending, first insertion ascending (Python's stable sort)
  int rank[2] = {0, 0};
  for (int j = 0; j < total; ++j) {
    const int owner = j & 31, slot = j >> 5;
    const bool hj = __shfl_sync(kFullMask, slot ? head[1] : head[0], owner);
    const double sj = __shfl_sync(kFullMask, slot ? acc[1] : acc[0], owner);
    if (!hj) continue;   // warp-uniform
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (head[e] && (sj > acc[e] || (sj == acc[e] && j < lane + 32 * e))) ++rank[e];
  }
  const int n_unique = __popc(__ballot_sync(kFullMask, head[0])) + __popc(__ballot_sync(kFullMask, head[1]));
  const int n_out = n_unique < top_n ? n_unique : top_n;
  const int64_t base = static_cast<int64_t>(q) * top_n;
#pragma unroll
  for (int e = 0; e < 2; ++e)
    if (head[e] && rank[e] < top_n) {
      out_ids[base + rank[e]] = id[e];
      out_scores[base + rank[e]] = acc[e];
    }
  for (int i = n_out + lane; i < top_n; i += 32) {
    out_ids[base + i] = -1;
    out_scores[base + i] = 0.0;
  }
  if (lane == 0 && out_counts) out_counts[q] = n_out;
}

__global__ void __launch_bounds__(kWrrfSmallWarps * 32)
wrrf_fuse_small_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ lens,
                       const double* __restrict__ weights, int n_lists, int list_stride,
                       double rrf_k, int top_n, int nq, int32_t* __restrict__ out_ids,
                       double* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kWrrfSmallWarps + (threadIdx.x >> 5);
  if (q >= nq) return;   // whole warp
  wrrf_small_warp(ids + static_cast<int64_t>(q) * n_lists * list_stride,
                  lens + static_cast<int64_t>(q) * n_lists, weights, n_lists, list_stride, rrf_k,
                  top_n, q, lane, out_ids, out_scores, out_counts);
}

// The hybrid step's two lists with the weights as kernel arguments: no device-side weight array to
// fill first (that was a 1-thread launch at the head of every step, in front of BOTH chains).
__global__ void __launch_bounds__(kWrrfSmallWarps * 32)
wrrf_fuse_pair_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ lens,
                      double w0, double w1, int list_stride, double rrf_k, int top_n, int nq,
                      int32_t* __restrict__ out_ids, double* __restrict__ out_scores,
                      int32_t* __restrict__ out_counts) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kWrrfSmallWarps + (threadIdx.x >> 5);
  if (q >= nq) return;   // whole warp
  const double weights[2] = {w0, w1};
  wrrf_small_warp(ids + static_cast<int64_t>(q) * 2 * list_stride, lens + static_cast<int64_t>(q) * 2,
                  weights, 2, list_stride, rrf_k, top_n, q, lane, out_ids, out_scores, out_counts);
}

bool wrrf_fuse_pair_fits(int list_stride) { return list_stride >= 1 && 2 * list_stride <= kWrrfSmallCap; }

cudaError_t launch_wrrf_fuse_pair(const int32_t* ids, const int32_t* lens, double w0, double w1,
                                  int list_stride, int nq, double rrf_k, int top_n, int32_t* out_ids,
                                  double* out_scores, int32_t* out_counts, cudaStream_t stream) {
  if (!wrrf_fuse_pair_fits(list_stride) || top_n < 1) return cudaErrorInvalidValue;
  if (nq < 1) return cudaSuccess;
  wrrf_fuse_pair_kernel<<<(nq + kWrrfSmallWarps - 1) / kWrrfSmallWarps, kWrrfSmallWarps * 32, 0,
                          stream>>>(ids, lens, w0, w1, list_stride, rrf_k, top_n, nq, out_ids,
                                    out_scores, out_counts);
  return cudaGetLastError();
}

// ---- the consumer of a sharded step's all-gather, in ONE launch ---------------------------------
// gathered: [n_parts][2][nq][k] sortable keys (per rank: dense plane, BM25 plane; ids global).  One
// warp per query merges each retriever's n_parts lists to its global top-k (rank by counting:
// keys are unique) in shared memory and runs the weighted RRF on the two merged lists -- what
// launch_topk_final x 2 + launch_wrrf_fuse do in four launches; at 8 GPUs those launches were a
// tenth of a 0.23 ms step.  Shapes: 2 k <= 64 and n_parts * k <= kShardMergeCap.
constexpr int kShardMergeCap = 256;

__global__ void __launch_bounds__(kWrrfSmallWarps * 32)
sharded_fuse_small_kernel(const uint64_t* __restrict__ gathered, int n_parts, int nq, int k,
                          double w_dense, double w_bm25, double rrf_k, int top_n,
                          int32_t* __restrict__ out_ids, double* __restrict__ out_scores,
                          int32_t* __restrict__ out_counts) {
  __shared__ uint64_t s_keys[kWrrfSmallWarps][kShardMergeCap];
  __shared__ int32_t s_ids[kWrrfSmallWarps][2][32];
  __shared__ int32_t s_lens[kWrrfSmallWarps][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kWrrfSmallWarps + warp;
  if (q >= nq) return;   // whole warp
  const int m = n_parts * k;
  uint64_t* keys = s_keys[warp];
  for (int which = 0; which < 2; ++which) {
    int live = 0;
    for (int i = lane; i < m; i += 32) {
      const int part = i / k, j = i - part * k;
      const uint64_t key =
          gathered[((static_cast<int64_t>(part) * 2 + which) * nq + q) * k + j];
      keys[i] = key;
      live += key != 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(kFullMask, live, o);
    __syncwarp();
    for (int i = lane; i < m; i += 32) {
      const uint64_t key = keys[i];
      if (key == 0ull) continue;
      int rank = 0;
      for (int j = 0; j < m; ++j) rank += keys[j] > key;
      if (rank < k) s_ids[warp][which][rank] = static_cast<int32_t>(key_id(key));
    }
    if (lane == 0) s_lens[warp][which] = live < k ? live : k;
    __syncwarp();
  }
  const double weights[2] = {w_dense, w_bm25};
  wrrf_small_warp(&s_ids[warp][0][0], s_lens[warp], weights, 2, 32, rrf_k, top_n, q, lane, out_ids,
                  out_scores, out_counts);
}

bool sharded_fuse_small_fits(int n_parts, int k) {
  return 2 * k <= kWrrfSmallCap && k <= 32 && n_parts * k <= kShardMergeCap;
}

cudaError_t launch_sharded_fuse_small(const uint64_t* gathered, int n_parts, int nq, int k,
                                      double w_dense, double w_bm25, double rrf_k, int top_n,
                                      int32_t* out_ids, double* out_scores, int32_t* out_counts,
                                      cudaStream_t stream) {
  if (!sharded_fuse_small_fits(n_parts, k) || nq < 1) return cudaErrorInvalidValue;
  sharded_fuse_small_kernel<<<(nq + kWrrfSmallWarps - 1) / kWrrfSmallWarps, kWrrfSmallWarps * 32, 0,
                              stream>>>(gathered, n_parts, nq, k, w_dense, w_bm25, rrf_k, top_n,
                                        out_ids, out_scores, out_counts);
  return cudaGetLastError();
}

size_t wrrf_scratch_keys(int n_lists, int list_stride, int nq) {
  const int64_t cap = static_cast<int64_t>(n_lists) * list_stride;
  if (cap <= kWrrfMaxEntries) return 0;
  return static_cast<size_t>(nq) * 2 * next_pow2(static_cast<int>(cap));
}

cudaError_t launch_wrrf_fuse(const int32_t* ids, const int32_t* lens, const double* weights,
                             int n_lists, int list_stride, int nq, double rrf_k, int top_n,
                             uint64_t* scratch, int32_t* out_ids, double* out_scores,
                             int32_t* out_counts, cudaStream_t stream) {
  if (n_lists < 1 || n_lists > 64 || list_stride < 1 || top_n < 1) return cudaErrorInvalidValue;
  const int64_t cap = static_cast<int64_t>(n_lists) * list_stride;
  if (cap > (1 << 22)) return cudaErrorInvalidConfiguration;
  if (cap <= kWrrfSmallCap) {
    if (nq < 1) return cudaSuccess;
    wrrf_fuse_small_kernel<<<(nq + kWrrfSmallWarps - 1) / kWrrfSmallWarps, kWrrfSmallWarps * 32, 0,
                             stream>>>(ids, lens, weights, n_lists, list_stride, rrf_k, top_n, nq,
                                       out_ids, out_scores, out_counts);
    return cudaGetLastError();
  }
  const int n_pow2 = next_pow2(static_cast<int>(cap) < 2 ? 2 : static_cast<int>(cap));
  const bool in_smem = cap <= kWrrfMaxEntries;
  if (!in_smem && !scratch) return cudaErrorInvalidValue;
  const int smem = in_smem ? n_pow2 * 16 : 0;
  cudaError_t e = cudaFuncSetAttribute(wrrf_fuse_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kWrrfMaxEntries * 16);
  if (e != cudaSuccess) return e;
  if (nq < 1) return cudaSuccess;
  wrrf_fuse_kernel<<<nq, kWrrfThreads, smem, stream>>>(ids, lens, weights, n_lists, list_stride,
                                                       rrf_k, top_n, n_pow2,
                                                       in_smem ? nullptr : scratch, out_ids,
                                                       out_scores, out_counts);
  return cudaGetLastError();
}

}  // namespace anr
