// BM25 over a CSR inverted index: gather each query term's postings, accumulate
// per-document scores in shared memory, select the top-k.
//
// Replaces bm25.get_scores(query_tokens) (rank_bm25.BM25Okapi, called at
// src/search_engine.py:219) + the top-k at :236-241 (and the filtered branch
// :221-234 through doc_mask).  Work decomposition: grid = (doc tiles, queries).
// One CTA owns a contiguous document range [d0, d1) whose fp32 accumulators live in
// shared memory.  Postings of a term are sorted by document, so the part of a
// term's posting list that falls in the tile is one contiguous slice, found by two
// binary searches; slices are streamed with coalesced loads.  A document occurs at
// most once per term, so within one term the scatter `acc[doc - d0] += idf * w` is
// conflict-free WITHOUT atomics; terms are processed one after another (query order,
// duplicates repeated -- the accumulation order of the reference loop) with a CTA
// barrier between them.  Tile top-k without per-row list maintenance: every thread takes the
// best key of its own documents, the k-th largest of those 256 thread-bests is a lower bound
// of the tile's k-th best (k distinct documents reach it), a second sweep collects the few
// documents at or above it (at most k threads can hold any) and one small sort ranks them.
// Zero-score documents are ordinary candidates (the reference returns them when fewer than k
// documents match); keys are unique (score, then lower document first), so ties are exact.
//
// Algorithmic HBM bytes per query: 8 B per posting of every query-term occurrence
// (4 B doc id + 4 B precomputed weight) + k*8 B of candidates per tile.
#include <cstdlib>

#include "anr_internal.h"
#include "anr_topk.cuh"

namespace anr {

constexpr int kBm25Threads = 256;
constexpr int kBm25TermChunk = 64;  // query terms whose slice bounds are staged at once

// weight[p] = tf*(k1+1) / (tf + k1*(1 - b + b*doc_len/avgdl)), evaluated in float64 in the
// operation order of BM25Okapi.get_scores, rounded once to fp32.
__global__ void __launch_bounds__(256)
bm25_weights_kernel(const int32_t* __restrict__ post_doc, const int32_t* __restrict__ post_tf,
                    const int32_t* __restrict__ doc_len, int64_t nnz, double k1, double b,
                    double avgdl, float* __restrict__ post_w) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < nnz;
       p += stride) {
    const double tf = static_cast<double>(post_tf[p]);
    const double dl = static_cast<double>(doc_len[post_doc[p]]);
    const double denom = __dadd_rn(tf, __dmul_rn(k1, __dadd_rn(__dadd_rn(1.0, -b),
                                                               __ddiv_rn(__dmul_rn(b, dl), avgdl))));
    post_w[p] = static_cast<float>(__ddiv_rn(__dmul_rn(tf, __dadd_rn(k1, 1.0)), denom));
  }
}

cudaError_t launch_bm25_weights(const int32_t* post_doc, const int32_t* post_tf,
                                const int32_t* doc_len, int64_t nnz, double k1, double b,
                                double avgdl, float* post_w, cudaStream_t stream) {
  if (nnz <= 0) return cudaSuccess;
  int64_t blocks = (nnz + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bm25_weights_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(post_doc, post_tf, doc_len,
                                                                         nnz, k1, b, avgdl, post_w);
  return cudaGetLastError();
}

Bm25Plan bm25_make_plan(const DeviceProps& dp, int n_docs, int nq, int k, bool emit_all) {
  Bm25Plan p;
  // enough (tile, query) CTAs for two waves over the SMs, tiles small enough that
  // several CTAs share an SM (48 KB of accumulators at most)
  int64_t tiles_wanted = (2LL * dp.sm_count + nq - 1) / nq;
  if (tiles_wanted < 1) tiles_wanted = 1;
  int64_t tile = (static_cast<int64_t>(n_docs) + tiles_wanted - 1) / tiles_wanted;
  int64_t tile_max = 6144;   // measured: 6144 beats 4096 / 8192 / 12288 / 24576 at batch 64 (profiles/)
  if (const char* e = getenv("ANR_BM25_TILE")) {   // tuning knob for profiling runs
    const int64_t v = atoll(e);
    if (v >= 1024 && v <= 49152) tile_max = v;
  }
  if (tile < 1024) tile = 1024;
  if (tile > tile_max) tile = tile_max;
  // the collect buffer holds k * ceil(tile / 256) keys at most: keep it <= 2048 entries
  if (!emit_all && k > 0) {
    const int64_t per_thread = 2048 / k > 1 ? 2048 / k : 1;
    if (tile > per_thread * kBm25Threads) tile = per_thread * kBm25Threads;
  }
  tile = (tile + 31) / 32 * 32;
  p.tile_docs = static_cast<int>(tile);
  p.n_tiles = n_docs > 0 ? static_cast<int>((n_docs + tile - 1) / tile) : 0;
  const int per_thread = (p.tile_docs + kBm25Threads - 1) / kBm25Threads;
  // [256 thread-bests][collect buffer: k * per_thread keys at most, a power of two for the sort]
  p.list_cap = emit_all ? 0
                        : kBm25Threads + next_pow2(k * per_thread < kBm25Threads ? kBm25Threads
                                                                                 : k * per_thread);
  p.smem_bytes = p.tile_docs * 4 + p.list_cap * 8 + kBm25TermChunk * (8 + 8 + 4);
  return p;
}

template <bool EMIT_ALL>
__global__ void __launch_bounds__(kBm25Threads)
bm25_score_kernel(Bm25View ix, const int32_t* __restrict__ q_terms,
                  const int32_t* __restrict__ q_offsets, int k,
                  const uint32_t* __restrict__ doc_mask, int tile_docs, int list_cap,
                  uint64_t* __restrict__ out, int64_t out_stride_q) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* acc = reinterpret_cast<float*>(smem);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(tile_docs) * 4);
  int64_t* s_lo = reinterpret_cast<int64_t*>(lists + list_cap);
  int64_t* s_hi = s_lo + kBm25TermChunk;
  float* s_idf = reinterpret_cast<float*>(s_hi + kBm25TermChunk);

  const int tile = blockIdx.x, q = blockIdx.y;
  const int d0 = tile * tile_docs;
  const int d1 = min(ix.n_docs, d0 + tile_docs);
  const int nd = d1 - d0;
  const int t_begin = q_offsets[q], t_end = q_offsets[q + 1];

  for (int i = threadIdx.x; i < nd; i += blockDim.x) acc[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = t_begin; c0 < t_end; c0 += kBm25TermChunk) {
    const int nc = min(kBm25TermChunk, t_end - c0);
    __syncthreads();  // previous chunk's bounds are no longer read; acc zeroing is visible
    // ---- first posting of every term inside the tile: one WARP per term, 32-ary search
    //      (5 dependent loads for 2^23 postings instead of 23 for a binary search) ----
    for (int j = warp; j < nc; j += kBm25Threads / 32) {
      const int term = q_terms[c0 + j];
      int64_t lo = 0, hi = 0;
      float idf = 0.f;
      if (term >= 0 && term < ix.n_terms) {
        idf = ix.idf[term];
        if (idf != 0.f) {  // `idf.get(q) or 0`: a zero idf contributes nothing
          lo = ix.term_ptr[term];
          hi = ix.term_ptr[term + 1];
        }
      }
      const int64_t term_end = hi;
      while (hi > lo) {   // invariant: the answer (first index with doc >= d0) lies in [lo, hi]
        const int64_t len = hi - lo;
        if (len <= 32) {
          const int64_t pos = lo + lane;
          const bool below = pos < hi && __ldg(ix.post_doc + pos) < d0;
          lo += __popc(__ballot_sync(kFullMask, below));
          break;
        }
        const int64_t step = (len + 32) / 33;
        const int64_t pos = lo + (lane + 1) * step - 1;
        const bool below = pos < hi && __ldg(ix.post_doc + pos) < d0;
        const int cnt = __popc(__ballot_sync(kFullMask, below));
        // probes 0..cnt-1 are below the target; probe cnt (if it exists and is in range) is not
        const int64_t nhi = lo + (cnt + 1) * step - 1;
        if (cnt < 32 && nhi < hi) hi = nhi;
        lo += cnt * step;
      }
      if (lane == 0) { s_lo[j] = lo; s_hi[j] = term_end; s_idf[j] = idf; }
    }
    __syncthreads();
    // ---- scatter-accumulate, one term after another.  A thread walks p = lo + tid, + 256, ...
    //      and stops at its first posting at or beyond the tile end (documents ascend), so no
    //      second search is needed; the first loads of term j+1 are issued BEFORE the barrier
    //      that closes term j, so their latency is not paid after it. ----
    struct Item { int d[4]; float w[4]; };
    const int kInvalid = 0x7fffffff;
    auto load_item = [&](int64_t p, int64_t end) {
      Item it;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t pp = p + u * kBm25Threads;
        const bool in = pp < end;
        it.d[u] = in ? __ldg(ix.post_doc + pp) : kInvalid;
        it.w[u] = in ? __ldg(ix.post_w + pp) : 0.f;
      }
      return it;
    };
    Item pre = load_item(s_lo[0] + threadIdx.x, s_hi[0]);
    for (int j = 0; j < nc; ++j) {
      const int64_t end = s_hi[j];
      const float idf = s_idf[j];
      int64_t p = s_lo[j] + threadIdx.x;
      Item cur = pre;
      while (true) {
        const bool more = cur.d[3] < d1;   // the 4th posting is the furthest: still inside?
        Item nxt;
        if (more) { p += 4 * kBm25Threads; nxt = load_item(p, end); }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cur.d[u] < d1) acc[cur.d[u] - d0] += idf * cur.w[u];
        if (!more) break;
        cur = nxt;
      }
      if (j + 1 < nc) pre = load_item(s_lo[j + 1] + threadIdx.x, s_hi[j + 1]);
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- emit / select ----
  if (EMIT_ALL) {
    for (int i = threadIdx.x; i < nd; i += blockDim.x) {
      const int doc = d0 + i;
      bool ok = true;
      if (doc_mask) ok = (__ldg(doc_mask + (doc >> 5)) >> (doc & 31)) & 1u;
      out[q * out_stride_q + doc] = ok ? make_key(acc[i], static_cast<uint32_t>(doc)) : 0ull;
    }
    return;
  }
  // ---- tile top-k ----
  __shared__ int n_sel;
  uint64_t* tbest = lists;                 // [256] thread-bests, sorted in place
  uint64_t* sel = lists + kBm25Threads;    // collected keys (list_cap - 256 slots)
  const int sel_cap = list_cap - kBm25Threads;
  auto key_of = [&](int i) -> uint64_t {
    const int doc = d0 + i;
    if (doc_mask && !((__ldg(doc_mask + (doc >> 5)) >> (doc & 31)) & 1u)) return 0ull;
    return make_key(acc[i], static_cast<uint32_t>(doc));
  };
  uint64_t best = 0ull;
  for (int i = threadIdx.x; i < nd; i += kBm25Threads) {
    const uint64_t key = key_of(i);
    best = key > best ? key : best;
  }
  if (threadIdx.x == 0) n_sel = 0;
  __shared__ uint64_t s_thr;
  if (k <= 32) {
    // k-th largest of the 256 thread-bests without a block-wide sort: every warp sorts its 32
    // keys in registers (shuffles, no barrier), warp 0 then pops the largest head k times
    uint64_t v = best;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const uint64_t other = __shfl_xor_sync(kFullMask, v, stride);
        const bool keep_max = ((lane & stride) == 0) == ((lane & size) == 0);
        v = keep_max ? (v > other ? v : other) : (v < other ? v : other);
      }
    tbest[threadIdx.x] = v;   // warp w's keys, descending, at tbest[32 w ..]
    __syncthreads();
    if (warp == 0) {
      int head = 0;   // lanes 0..7: read position in warp `lane`'s sorted run
      uint64_t kth = 0ull;
      for (int it = 0; it < k; ++it) {
        const uint64_t c = (lane < kBm25Threads / 32 && head < 32) ? tbest[lane * 32 + head] : 0ull;
        uint64_t m = c;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          const uint64_t other = __shfl_xor_sync(kFullMask, m, o);
          m = other > m ? other : m;
        }
        m = __shfl_sync(kFullMask, m, 0);   // max over lanes 0..7
        kth = m;
        if (m == 0ull) break;               // fewer than k threads hold a document
        if (c == m) ++head;                 // keys are unique: exactly one lane advances
      }
      if (lane == 0) s_thr = kth;
    }
    __syncthreads();
  } else {
    tbest[threadIdx.x] = best;
    block_bitonic_sort_desc(tbest, kBm25Threads);
    if (threadIdx.x == 0) s_thr = k <= kBm25Threads ? tbest[k - 1] : 0ull;
    __syncthreads();
  }
  const uint64_t thr = s_thr;   // 0: fewer than k threads hold a document -> everything passes
  for (int i = threadIdx.x; i < sel_cap; i += kBm25Threads) sel[i] = 0ull;
  __syncthreads();
  for (int i = threadIdx.x; i < nd; i += kBm25Threads) {
    const uint64_t key = key_of(i);
    if (key != 0ull && key >= thr) {
      const int slot = atomicAdd(&n_sel, 1);
      if (slot < sel_cap) sel[slot] = key;   // cannot overflow: <= k threads x ceil(tile / 256) docs
    }
  }
  __syncthreads();
  const int ns = n_sel < sel_cap ? n_sel : sel_cap;
  uint64_t* o = out + q * out_stride_q + static_cast<int64_t>(tile) * k;
  if (ns <= kBm25Threads) {
    // rank by counting (keys are unique): one pass, no sort
    for (int i = threadIdx.x; i < k; i += kBm25Threads)
      if (i >= ns) o[i] = 0ull;
    if (threadIdx.x < ns) {
      const uint64_t key = sel[threadIdx.x];
      int rank = 0;
      for (int j = 0; j < ns; ++j) rank += sel[j] > key;
      if (rank < k) o[rank] = key;
    }
  } else {
    block_bitonic_sort_desc(sel, next_pow2(ns));
    for (int i = threadIdx.x; i < k; i += blockDim.x) o[i] = i < ns ? sel[i] : 0ull;
  }
}

template <bool EMIT_ALL>
static cudaError_t launch_score_t(const Bm25View& ix, const int32_t* q_terms,
                                  const int32_t* q_offsets, int nq, int k,
                                  const uint32_t* doc_mask, const Bm25Plan& plan, uint64_t* out,
                                  int64_t out_stride_q, cudaStream_t stream) {
  if (plan.n_tiles < 1 || nq < 1) return cudaSuccess;
  auto kern = bm25_score_kernel<EMIT_ALL>;
  cudaError_t e =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
  if (e != cudaSuccess) return e;
  // grid.y is limited to 65535 queries per launch
  for (int q0 = 0; q0 < nq; q0 += 65535) {
    const int nb = nq - q0 < 65535 ? nq - q0 : 65535;
    dim3 grid(plan.n_tiles, nb);
    kern<<<grid, kBm25Threads, plan.smem_bytes, stream>>>(ix, q_terms, q_offsets + q0, k, doc_mask,
                                                          plan.tile_docs, plan.list_cap,
                                                          out + q0 * out_stride_q, out_stride_q);
  }
  return cudaGetLastError();
}

cudaError_t launch_bm25_score_topk(const Bm25View& ix, const int32_t* q_terms,
                                   const int32_t* q_offsets, int nq, int k,
                                   const uint32_t* doc_mask, const Bm25Plan& plan, uint64_t* cand,
                                   int64_t cand_stride_q, cudaStream_t stream) {
  if (k < 1 || k > kMaxFusedK) return cudaErrorInvalidValue;
  return launch_score_t<false>(ix, q_terms, q_offsets, nq, k, doc_mask, plan, cand, cand_stride_q,
                               stream);
}

cudaError_t launch_bm25_score_all(const Bm25View& ix, const int32_t* q_terms,
                                  const int32_t* q_offsets, int nq, const uint32_t* doc_mask,
                                  const Bm25Plan& plan, uint64_t* keys, int64_t keys_stride_q,
                                  cudaStream_t stream) {
  return launch_score_t<true>(ix, q_terms, q_offsets, nq, 1, doc_mask, plan, keys, keys_stride_q,
                              stream);
}

// Raw fp32 scores of every document for ONE query (BM25Okapi.get_scores itself).
__global__ void __launch_bounds__(256)
keys_to_scores_kernel(const uint64_t* __restrict__ keys, int64_t n, float* __restrict__ scores) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) scores[i] = keys[i] ? key_score(keys[i]) : 0.f;
}

cudaError_t launch_keys_to_scores(const uint64_t* keys, int64_t n, float* scores,
                                  cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  keys_to_scores_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(keys, n, scores);
  return cudaGetLastError();
}

}  // namespace anr
