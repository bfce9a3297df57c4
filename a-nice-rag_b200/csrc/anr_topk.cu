// Final top-k selection over candidate keys, and the large-k (full ranking) sort path.
//
// Small k (<= kMaxFusedK): one CTA per query; if the candidates fit in shared memory they
// are bitonic-sorted directly, otherwise 32 warps first reduce them with threshold lists.
// Large k (the evaluator's similarity_k = 12000 >= N case, src/retrieval_eval.py:142 ->
// the `else: argsort()[::-1]` branches of src/search_engine.py:86-87,134-135,240-241):
// all keys are materialised and sorted descending by a global-memory bitonic network whose
// inner strides run in shared memory.
#include "anr_internal.h"
#include "anr_topk.cuh"

namespace anr {

// Candidate i of a query lives at base[(i / seg_len) * seg_stride + (i % seg_len)]:
// contiguous candidates use seg_len = m; the all-gathered [part][query][k] layout of
// the sharded search uses seg_len = k, seg_stride = n_queries * k.
__global__ void __launch_bounds__(kFinalThreads)
topk_final_kernel(const uint64_t* __restrict__ cand, int64_t stride_q, int m, int seg_len,
                  int64_t seg_stride, int k, TopkOut o, const int32_t* __restrict__ n_active,
                  const int32_t* __restrict__ row_map) {
  __shared__ uint64_t keys[kFinalSortCap];
  pdl_wait();   // (no early trigger: the last kernel of a chain; what follows it -- the fusion, an
                // event record, a copy to the host -- must see it complete)
  if (n_active && static_cast<int>(blockIdx.x) >= *n_active) return;
  const uint64_t* c = cand + blockIdx.x * stride_q;
  const int q = row_map ? row_map[blockIdx.x] : blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int p2;
  if (m <= kFinalSortCap) {
    p2 = next_pow2(m);
    for (int i = threadIdx.x; i < p2; i += blockDim.x)
      keys[i] = i < m ? c[(i / seg_len) * seg_stride + (i % seg_len)] : 0ull;
  } else {
    const int n_lists = kFinalThreads / 32;
    p2 = next_pow2(n_lists * k);
    for (int i = threadIdx.x; i < p2; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    uint64_t thr = 0;
    uint64_t* list = keys + warp * k;
    for (int base = warp * 32; base < m; base += kFinalThreads) {
      const int i = base + lane;
      const uint64_t key = i < m ? c[(i / seg_len) * seg_stride + (i % seg_len)] : 0ull;
      warp_list_offer(list, k, key, thr, lane);
    }
  }
  block_bitonic_sort_desc(keys, p2);
  const bool dead = o.q_offsets && o.q_offsets[q + 1] == o.q_offsets[q];   // a query without terms
  uint64_t key = 0ull;
  if (threadIdx.x < k) {
    key = (threadIdx.x < p2 && !dead) ? keys[threadIdx.x] : 0ull;
    emit_entry(key, q * o.stride_q + threadIdx.x, o);
  }
  const int cnt = __syncthreads_count(key != 0ull);
  if (threadIdx.x == 0 && o.counts) o.counts[q * o.count_stride] = cnt;
}

// Small candidate sets (k * ceil(m / 256) <= kSmallCap): the k-th largest of the 256 per-thread
// bests bounds the k-th best from below; one sweep collects the few keys at or above it and one
// short sort ranks them -- an order of magnitude fewer barriers than sorting all m keys.
constexpr int kSmallThreads = 256;
constexpr int kSmallCap = 2048;

// n_active / row_map (nullable): only the first *n_active blocks work, block b writes result
// row row_map[b] (the device-side list of queries the tensor-core path flagged).
__global__ void __launch_bounds__(kSmallThreads)
topk_final_small_kernel(const uint64_t* __restrict__ cand, int64_t stride_q, int m, int seg_len,
                        int64_t seg_stride, int k, TopkOut o,
                        const int32_t* __restrict__ n_active, const int32_t* __restrict__ row_map) {
  __shared__ uint64_t best[kSmallThreads];
  __shared__ uint64_t sel[kSmallCap];
  __shared__ int n_sel;
  pdl_wait();   // (no early trigger: the last kernel of a chain; what follows it -- the fusion, an
                // event record, a copy to the host -- must see it complete)
  if (n_active && static_cast<int>(blockIdx.x) >= *n_active) return;
  const uint64_t* c = cand + blockIdx.x * stride_q;
  const int q = row_map ? row_map[blockIdx.x] : blockIdx.x;
  uint64_t b = 0ull;
  for (int i = threadIdx.x; i < m; i += kSmallThreads) {
    const uint64_t v = c[(i / seg_len) * seg_stride + (i % seg_len)];
    b = v > b ? v : b;
  }
  best[threadIdx.x] = b;
  if (threadIdx.x == 0) n_sel = 0;
  block_bitonic_sort_desc(best, kSmallThreads);
  const uint64_t thr = k <= kSmallThreads ? best[k - 1] : 0ull;
  const int want = k * ((m + kSmallThreads - 1) / kSmallThreads);
  const int cap = next_pow2(want < 2 ? 2 : want);
  for (int i = threadIdx.x; i < cap; i += kSmallThreads) sel[i] = 0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += kSmallThreads) {
    const uint64_t v = c[(i / seg_len) * seg_stride + (i % seg_len)];
    if (v != 0ull && v >= thr) {
      const int slot = atomicAdd(&n_sel, 1);
      if (slot < cap) sel[slot] = v;
    }
  }
  __syncthreads();
  const int ns = n_sel < cap ? n_sel : cap;
  block_bitonic_sort_desc(sel, next_pow2(ns < 2 ? 2 : ns));
  const bool dead = o.q_offsets && o.q_offsets[q + 1] == o.q_offsets[q];   // a query without terms
  uint64_t key = 0ull;
  for (int i = threadIdx.x; i < k; i += kSmallThreads) {
    key = (i < ns && !dead) ? sel[i] : 0ull;
    emit_entry(key, q * o.stride_q + i, o);
  }
  if (threadIdx.x == 0 && o.counts) o.counts[q * o.count_stride] = dead ? 0 : (ns < k ? ns : k);
}

cudaError_t launch_topk_final(const uint64_t* cand, int64_t cand_stride_q, int m, int seg_len,
                              int64_t seg_stride, int nq, int k, const TopkOut& out,
                              cudaStream_t stream) {
  if (k < 1 || k > kMaxFusedK || m < 1 || seg_len < 1) return cudaErrorInvalidValue;
  if (nq < 1) return cudaSuccess;
  if (static_cast<int64_t>(k) * ((m + kSmallThreads - 1) / kSmallThreads) <= kSmallCap) {
    topk_final_small_kernel<<<nq, kSmallThreads, 0, stream>>>(cand, cand_stride_q, m, seg_len,
                                                              seg_stride, k, out, nullptr, nullptr);
    return cudaGetLastError();
  }
  topk_final_kernel<<<nq, kFinalThreads, 0, stream>>>(cand, cand_stride_q, m, seg_len, seg_stride,
                                                      k, out, nullptr, nullptr);
  return cudaGetLastError();
}

// flags[0..nq) -> n_flagged, flagged[] (ascending query indices); one small block
__global__ void __launch_bounds__(256)
compact_flags_kernel(const int32_t* __restrict__ flags, int nq, int32_t* __restrict__ n_flagged,
                     int32_t* __restrict__ flagged) {
  __shared__ int count;
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  for (int base = 0; base < nq; base += 256) {   // ordered: one chunk of 256 queries at a time
    const int q = base + threadIdx.x;
    const bool f = q < nq && flags[q] != 0;
    const unsigned ballot = __ballot_sync(kFullMask, f);
    __shared__ int warp_off[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) warp_off[warp] = __popc(ballot);
    __syncthreads();
    int off = count;
    for (int w = 0; w < warp; ++w) off += warp_off[w];
    if (f) flagged[off + __popc(ballot & ((1u << lane) - 1))] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
      int total = 0;
      for (int w = 0; w < 8; ++w) total += warp_off[w];
      count += total;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_flagged = count;
}

cudaError_t launch_compact_flags(const int32_t* flags, int nq, int32_t* n_flagged, int32_t* flagged,
                                 cudaStream_t stream) {
  return launch_chain(compact_flags_kernel, dim3(1), dim3(256), 0, stream, flags, nq, n_flagged,
                      flagged);
}

// Merge of the flagged rescan: block b < *n_flagged ranks cand[b] into result row flagged[b].
cudaError_t launch_topk_final_flagged(const uint64_t* cand, int64_t cand_stride, int m, int nq,
                                      int k, const TopkOut& out, const int32_t* n_flagged,
                                      const int32_t* flagged, cudaStream_t stream) {
  if (k < 1 || k > kMaxFusedK || m < 1) return cudaErrorInvalidValue;
  // (these launches usually find an empty list and run beside another pass' main kernel: the same
  // L1 / shared-memory split as the dense kernels, so that their CTAs can share an SM with it)
  static const bool carveout_set = [] {
    cudaFuncSetAttribute(topk_final_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(topk_final_small_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  (void)carveout_set;
  if (static_cast<int64_t>(k) * ((m + kSmallThreads - 1) / kSmallThreads) > kSmallCap)
    return launch_chain(topk_final_kernel, dim3(nq), dim3(kFinalThreads), 0, stream, cand, cand_stride,
                        m, m, 0, k, out, n_flagged, flagged);
  return launch_chain(topk_final_small_kernel, dim3(nq), dim3(kSmallThreads), 0, stream, cand,
                      cand_stride, m, m, 0, k, out, n_flagged, flagged);
}

// ---- large-k path: global bitonic sort, descending --------------------------------
constexpr int kSortTile = 2048;  // keys per CTA in the shared-memory phases
constexpr int kSortThreads = 1024;

__device__ __forceinline__ void cmp_swap_desc(uint64_t& a, uint64_t& b, bool desc) {
  if ((a < b) == desc) { uint64_t t = a; a = b; b = t; }
}

// All network steps with size <= kSortTile, entirely in shared memory.
__global__ void __launch_bounds__(kSortThreads)
bitonic_head_kernel(uint64_t* __restrict__ keys, int64_t stride_q, int64_t n_pow2) {
  __shared__ uint64_t s[kSortTile];
  uint64_t* base = keys + blockIdx.y * stride_q + static_cast<int64_t>(blockIdx.x) * kSortTile;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * kSortTile;
  const int tile = n_pow2 < kSortTile ? static_cast<int>(n_pow2) : kSortTile;
  for (int i = threadIdx.x; i < tile; i += blockDim.x) s[i] = base[i];
  for (int size = 2; size <= tile; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (tile >> 1); t += blockDim.x) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const bool desc = (((g0 + lo) & size) == 0);
        cmp_swap_desc(s[lo], s[lo + stride], desc);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tile; i += blockDim.x) base[i] = s[i];
}

// One (size, stride) step with stride >= kSortTile, in global memory.
__global__ void __launch_bounds__(256)
bitonic_global_step_kernel(uint64_t* __restrict__ keys, int64_t stride_q, int64_t n_pow2,
                           int64_t size, int64_t stride) {
  uint64_t* base = keys + blockIdx.y * stride_q;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= (n_pow2 >> 1)) return;
  const int64_t lo = ((t / stride) * (stride << 1)) + (t % stride);
  const bool desc = ((lo & size) == 0);
  uint64_t a = base[lo], b = base[lo + stride];
  if ((a < b) == desc) { base[lo] = b; base[lo + stride] = a; }
}

// The strides < kSortTile of one `size` (> kSortTile), in shared memory.
__global__ void __launch_bounds__(kSortThreads)
bitonic_tail_kernel(uint64_t* __restrict__ keys, int64_t stride_q, int64_t size) {
  __shared__ uint64_t s[kSortTile];
  uint64_t* base = keys + blockIdx.y * stride_q + static_cast<int64_t>(blockIdx.x) * kSortTile;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * kSortTile;
  for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) s[i] = base[i];
  const bool desc = ((g0 & size) == 0);  // constant inside a tile since size > kSortTile
  for (int stride = kSortTile >> 1; stride > 0; stride >>= 1) {
    __syncthreads();
    for (int t = threadIdx.x; t < (kSortTile >> 1); t += blockDim.x) {
      const int lo = ((t / stride) * (stride << 1)) + (t % stride);
      cmp_swap_desc(s[lo], s[lo + stride], desc);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) base[i] = s[i];
}

cudaError_t launch_sort_desc(uint64_t* keys, int64_t stride_q, int64_t n_pow2, int nq,
                             cudaStream_t stream) {
  if (n_pow2 < 1 || (n_pow2 & (n_pow2 - 1)) != 0) return cudaErrorInvalidValue;
  if (n_pow2 == 1) return cudaSuccess;
  const int64_t tiles = n_pow2 <= kSortTile ? 1 : n_pow2 / kSortTile;
  dim3 grid_tiles(static_cast<unsigned>(tiles), nq);
  bitonic_head_kernel<<<grid_tiles, kSortThreads, 0, stream>>>(keys, stride_q, n_pow2);
  for (int64_t size = static_cast<int64_t>(kSortTile) * 2; size <= n_pow2; size <<= 1) {
    for (int64_t stride = size >> 1; stride >= kSortTile; stride >>= 1) {
      dim3 grid(static_cast<unsigned>(((n_pow2 >> 1) + 255) / 256), nq);
      bitonic_global_step_kernel<<<grid, 256, 0, stream>>>(keys, stride_q, n_pow2, size, stride);
    }
    bitonic_tail_kernel<<<grid_tiles, kSortThreads, 0, stream>>>(keys, stride_q, size);
  }
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
emit_sorted_kernel(const uint64_t* __restrict__ keys, int64_t stride_q, int64_t n_avail, int k,
                   TopkOut o) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int q = blockIdx.x;
  const bool dead = o.q_offsets && o.q_offsets[q + 1] == o.q_offsets[q];   // a query without terms
  int mine = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = (i < n_avail && !dead) ? keys[q * stride_q + i] : 0ull;
    emit_entry(key, q * o.stride_q + i, o);
    mine += key != 0ull;
  }
  if (mine) atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0 && o.counts) o.counts[q * o.count_stride] = cnt;
}

// keys[q][0 .. n_avail) sorted descending -> first k entries per query
cudaError_t launch_emit_sorted(const uint64_t* keys, int64_t stride_q, int64_t n_avail, int nq,
                               int k, const TopkOut& out, cudaStream_t stream) {
  if (nq < 1) return cudaSuccess;
  emit_sorted_kernel<<<nq, 256, 0, stream>>>(keys, stride_q, n_avail, k, out);
  return cudaGetLastError();
}

// zero keys[q][n .. n_pow2) so the padding sorts last
__global__ void __launch_bounds__(256)
zero_tail_kernel(uint64_t* __restrict__ keys, int64_t stride_q, int64_t n, int64_t n_pow2) {
  uint64_t* base = keys + blockIdx.y * stride_q;
  for (int64_t i = n + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pow2;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    base[i] = 0ull;
}

cudaError_t launch_zero_tail(uint64_t* keys, int64_t stride_q, int64_t n, int64_t n_pow2, int nq,
                             cudaStream_t stream) {
  if (n >= n_pow2 || nq < 1) return cudaSuccess;
  int64_t blocks = (n_pow2 - n + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  dim3 grid(static_cast<unsigned>(blocks), nq);
  zero_tail_kernel<<<grid, 256, 0, stream>>>(keys, stride_q, n, n_pow2);
  return cudaGetLastError();
}

}  // namespace anr
