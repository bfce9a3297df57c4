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
  if (tile < 1024) tile = 1024;
  if (tile > 12288) tile = 12288;
  // the collect buffer holds k * ceil(tile / 256) keys at most: keep it <= 2048 entries
  if (!emit_all && k > 0) {
    const int64_t per_thread = 2048 / k > 1 ? 2048 / k : 1;
    if (tile > per_thread * kBm25Threads) tile = per_thread * kBm25Threads;
  }
  tile = (tile + 31) / 32 * 32;
  p.tile_docs = static_cast<int>(tile);
  p.n_tiles = n_docs > 0 ? static_cast<int>((n_docs + tile - 1) / tile) : 0;
  const int per_thread = (p.tile_docs + kBm25Threads - 1) / kBm25Threads;
  p.list_cap = emit_all ? 0 : next_pow2(k * per_thread < 2 * kBm25Threads ? 2 * kBm25Threads
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

  for (int c0 = t_begin; c0 < t_end; c0 += kBm25TermChunk) {
    const int nc = min(kBm25TermChunk, t_end - c0);
    __syncthreads();  // previous chunk's bounds are no longer read; acc zeroing is visible
    // ---- slice bounds: thread j -> lower bound of d0, thread nc + j -> lower bound of d1 ----
    if (threadIdx.x < 2 * nc) {
      const int j = threadIdx.x < nc ? threadIdx.x : threadIdx.x - nc;
      const int target = threadIdx.x < nc ? d0 : d1;
      const int term = q_terms[c0 + j];
      int64_t lo = 0, hi = 0;
      float idf = 0.f;
      if (term >= 0 && term < ix.n_terms) {
        idf = ix.idf[term];
        if (idf != 0.f) {  // `idf.get(q) or 0`: a zero idf contributes nothing
          lo = ix.term_ptr[term];
          hi = ix.term_ptr[term + 1];
          while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(ix.post_doc + mid) < target) lo = mid + 1; else hi = mid;
          }
        }
      }
      if (threadIdx.x < nc) { s_lo[j] = lo; s_idf[j] = idf; } else { s_hi[j] = lo; }
    }
    __syncthreads();
    // ---- scatter-accumulate, one term after another ----
    for (int j = 0; j < nc; ++j) {
      const int64_t lo = s_lo[j], hi = s_hi[j];
      const float idf = s_idf[j];
      int64_t p = lo + threadIdx.x;
      // 4 postings in flight per thread
      for (; p + 3 * kBm25Threads < hi; p += 4 * kBm25Threads) {
        int dd[4];
        float ww[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dd[u] = __ldg(ix.post_doc + p + u * kBm25Threads);
          ww[u] = __ldg(ix.post_w + p + u * kBm25Threads);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[dd[u] - d0] += idf * ww[u];
      }
      for (; p < hi; p += kBm25Threads)
        acc[__ldg(ix.post_doc + p) - d0] += idf * __ldg(ix.post_w + p);
      if (lo < hi) __syncthreads();  // lo/hi are CTA-uniform
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
  tbest[threadIdx.x] = best;
  if (threadIdx.x == 0) n_sel = 0;
  block_bitonic_sort_desc(tbest, kBm25Threads);
  const uint64_t thr = k <= kBm25Threads ? tbest[k - 1] : 0ull;   // 0: fewer than k threads hold a document
  __syncthreads();
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
  block_bitonic_sort_desc(sel, next_pow2(ns < 2 ? 2 : ns));
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    out[q * out_stride_q + static_cast<int64_t>(tile) * k + i] = i < ns ? sel[i] : 0ull;
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
