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
//
// PRUNE = true (top-k only): safe dynamic pruning in the spirit of MaxScore.  Under a Zipfian
// vocabulary ~85 % of a query's postings belong to a few frequent, low-idf HEAD terms.  Those are
// held a second time as dense rows head_w[slot][doc], so the kernel streams only the postings of
// the other terms (the "essential" ones), takes the tile's k-th best PARTIAL score theta (a
// lower bound of its k-th best full score, head contributions being non-negative; raised further
// by the best theta any earlier tile of the same query has published), and completes with one
// load per head term only the documents whose partial score plus an upper bound of the head
// contributions can still reach theta.  Everything else provably lies below the tile's top-k.
#include <cstdlib>

#include "anr_internal.h"
#include "anr_topk.cuh"

namespace anr {

constexpr int kBm25Threads = 256;
constexpr int kBm25TermChunk = 64;  // query terms whose slice bounds are staged at once
constexpr int kBm25CandCap = 2048;  // pruned scan: documents of a tile whose score is completed

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
  p.smem_bytes = p.tile_docs * 4 + p.list_cap * 8 + kBm25TermChunk * (8 + 8 + 4) +
                 (emit_all ? 0 : kBm25CandCap * 2);
  return p;
}

__device__ __forceinline__ uint64_t kth_of_thread_bests(uint64_t best, int k, uint64_t* tbest,
                                                        uint64_t* s_out) {
  return block_kth_of_thread_bests<kBm25Threads / 32>(best, k, tbest, s_out);
}

// One (tile, query) work item; bx / by = its block coordinates in the launch grid.
template <bool EMIT_ALL, bool PRUNE>
__device__ __forceinline__ void
bm25_tile_item(const Bm25View& ix, const Bm25HeadView& hd, const int32_t* __restrict__ q_terms,
               const int32_t* __restrict__ q_offsets, int k, const uint32_t* __restrict__ doc_mask,
               int tile_docs, int list_cap, uint64_t* __restrict__ out, int64_t out_stride_q,
               float* __restrict__ theta_g, int tile_stride, int n_sampled, int bx, int by,
               const int32_t* __restrict__ q_list = nullptr) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* acc = reinterpret_cast<float*>(smem);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(tile_docs) * 4);
  int64_t* s_lo = reinterpret_cast<int64_t*>(lists + list_cap);
  int64_t* s_hi = s_lo + kBm25TermChunk;
  float* s_idf = reinterpret_cast<float*>(s_hi + kBm25TermChunk);
  uint16_t* cand_idx = reinterpret_cast<uint16_t*>(s_idf + kBm25TermChunk);   // [kBm25CandCap]
  __shared__ int n_stream, n_cand;
  // pruned scan: head terms of the query (slot, idf) in query order, bound of their contribution
  __shared__ int s_slot[kBm25TermChunk];
  __shared__ int h_slot[kBm25TermChunk];
  __shared__ float h_idf[kBm25TermChunk];
  __shared__ int n_h;
  __shared__ float h_ub;

  // n_sampled > 0: grid = (queries, tiles), queries fastest, and the tiles are visited in an order
  // that starts with every tile_stride-th one: the first wave of CTAs plays the role of a sample
  // pass (it publishes a k-th best score per query) and everything dispatched later prunes
  // against it.  n_sampled <= 0: grid = (tiles, queries), natural order.
  int tile, q;
  if (n_sampled == -2) {   // separate sample launch: every tile_stride-th tile
    tile = bx * tile_stride;
    q = by;
  } else if (n_sampled < -2) {    // main launch after a separate sample launch: skip its tiles
    tile = bx;
    q = by;
    if (tile % tile_stride == 0) return;
  } else if (n_sampled > 0) {
    q = bx;
    const int y = by;
    if (y < n_sampled) {
      tile = y * tile_stride;
    } else {
      const int j = y - n_sampled;
      tile = j + j / (tile_stride - 1) + 1;   // the j-th tile that is not a multiple of tile_stride
    }
  } else {
    tile = bx;
    q = by;
  }
  // q_list (natural order only): block row `by` scores query q_list[by] into output row `by`
  const int out_row = q;
  if (q_list) q = q_list[by];
  const int d0 = tile * tile_docs;
  const int d1 = min(ix.n_docs, d0 + tile_docs);
  const int nd = d1 - d0;
  const int t_begin = q_offsets[q], t_end = q_offsets[q + 1];

  {  // tile_docs is a multiple of 32: whole float4s
    float4* a4 = reinterpret_cast<float4*>(acc);
    const int n4 = (nd + 3) >> 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += kBm25Threads) a4[i] = z;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a query of more than one staged chunk of terms takes the unpruned path
  const bool prune = PRUNE && hd.n_head > 0 && (t_end - t_begin) <= kBm25TermChunk;
  if (PRUNE && threadIdx.x == 0) { n_h = 0; h_ub = 0.f; }
  for (int c0 = t_begin; c0 < t_end; c0 += kBm25TermChunk) {
    const int nc = min(kBm25TermChunk, t_end - c0);
    __syncthreads();  // previous chunk's bounds are no longer read; acc zeroing is visible
    if (prune) {
      // ---- which head terms are left to the dense rows (non-essential)?  Warp 0, one lane per
      // term (two for 33..64 terms).  With a published bound theta for this query, only as many
      // head terms -- cheapest upper bound first -- as keep the sum of their bounds BELOW theta:
      // then the pruning test always has room (cut = theta - ub > 0) and no tile falls back to
      // completing every document (queries with 6-7 head terms did: 19 % of all instructions).
      // The other head terms are streamed from their postings like any tail term.  Without a
      // bound (sample tiles) every head term with a positive idf is non-essential, as before.
      if (warp == 0) {
        const float theta0 = theta_g ? __ldcg(theta_g + q) : 0.f;
        int slot_e[2] = {-1, -1};
        float ub_e[2] = {0.f, 0.f};
        float tail_idf = 0.f;   // largest idf among my terms that are not head terms
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = lane + 32 * e;
          if (j < nc) {
            const int term = q_terms[c0 + j];
            if (term >= 0 && term < ix.n_terms) {
              const float idf = ix.idf[term];
              const int sl = hd.slot_of[term];
              if (idf > 0.f && sl != 0xff) {
                slot_e[e] = sl;
                ub_e[e] = idf * hd.head_max[sl];
              } else {
                tail_idf = fmaxf(tail_idf, idf);
              }
            }
          }
        }
        // No bound published yet (sample tiles): GUESS one for the classification only -- the idf
        // of the rarest other term (a document that holds it typically scores at least that).  A
        // wrong guess costs time, never correctness: the pruning test uses the real bounds.
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tail_idf = fmaxf(tail_idf, __shfl_xor_sync(kFullMask, tail_idf, o));
        const float theta_cls = theta0 > 0.f ? theta0 : tail_idf;
        if (theta_cls > 0.f) {
          // cum[e] = sum of the bounds of all head occurrences ordered at or before mine by (ub, j)
          float cum[2] = {0.f, 0.f};
          for (int jj = 0; jj < nc; ++jj) {
            const int owner = jj & 31, sl2 = jj >> 5;
            const int s_jj = __shfl_sync(kFullMask, sl2 ? slot_e[1] : slot_e[0], owner);
            const float u_jj = __shfl_sync(kFullMask, sl2 ? ub_e[1] : ub_e[0], owner);
            if (s_jj < 0) continue;   // warp-uniform
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = lane + 32 * e;
              if (slot_e[e] >= 0 && (u_jj < ub_e[e] || (u_jj == ub_e[e] && jj <= j))) cum[e] += u_jj;
            }
          }
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (slot_e[e] >= 0 && !(cum[e] * (1.f + 1e-4f) < theta_cls)) slot_e[e] = -1;   // essential
        }
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (lane + 32 * e < nc) s_slot[lane + 32 * e] = slot_e[e];
      }
      __syncthreads();
    }
    // ---- first posting of every term inside the tile: one WARP per term, 32-ary search
    //      (5 dependent loads for 2^23 postings instead of 23 for a binary search) ----
    for (int j = warp; j < nc; j += kBm25Threads / 32) {
      const int term = q_terms[c0 + j];
      int64_t lo = 0, hi = 0;
      float idf = 0.f;
      const int slot = prune ? s_slot[j] : -1;
      if (term >= 0 && term < ix.n_terms) {
        idf = ix.idf[term];
        // `idf.get(q) or 0`: a zero idf contributes nothing; a non-essential head term is NOT
        // streamed by the pruned scan
        if (idf != 0.f && slot < 0) {
          lo = ix.term_ptr[term];
          hi = ix.term_ptr[term + 1];
        }
      }
      if (PRUNE && !prune && lane == 0) s_slot[j] = -1;
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
    if (prune && threadIdx.x == 0) {   // head terms in query order (deterministic summation order)
      int n = 0;
      float ub = 0.f;
      for (int j = 0; j < nc; ++j)
        if (s_slot[j] >= 0) {
          h_slot[n] = s_slot[j];
          h_idf[n] = s_idf[j];
          ub += s_idf[j] * hd.head_max[s_slot[j]];
          ++n;
        }
      n_h = n;
      h_ub = ub;
    }
    if (threadIdx.x == 0) {   // only terms with postings to stream take a turn (and a barrier) below
      int m = 0;
      for (int j = 0; j < nc; ++j)
        if (s_hi[j] > s_lo[j]) {
          s_lo[m] = s_lo[j];
          s_hi[m] = s_hi[j];
          s_idf[m] = s_idf[j];
          ++m;
        }
      n_stream = m;
    }
    __syncthreads();
    const int ns_terms = n_stream;
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
    for (int j = 0; j < ns_terms; ++j) {
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
      if (j + 1 < ns_terms) pre = load_item(s_lo[j + 1] + threadIdx.x, s_hi[j + 1]);
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
      out[out_row * out_stride_q + doc] = ok ? make_key(acc[i], static_cast<uint32_t>(doc)) : 0ull;
    }
    return;
  }
  // ---- tile top-k ----
  __shared__ int n_sel;
  __shared__ uint64_t s_thr;
  uint64_t* tbest = lists;                 // [256] thread-bests, sorted in place
  uint64_t* sel = lists + kBm25Threads;    // collected keys (list_cap - 256 slots)
  const int sel_cap = list_cap - kBm25Threads;
  // Pruned scan: `cut` excludes every document whose accumulator lies below it (completing a
  // document only raises its accumulator, so the test holds before and after completion).
  float cut = -INFINITY;
  auto key_of = [&](int i) -> uint64_t {
    const int doc = d0 + i;
    if (doc_mask && !((__ldg(doc_mask + (doc >> 5)) >> (doc & 31)) & 1u)) return 0ull;
    const float a = acc[i];
    if (PRUNE && a < cut) return 0ull;
    return make_key(a, static_cast<uint32_t>(doc));
  };
  uint64_t best = 0ull;
  bool listed = false;      // survivors are exactly cand_idx[0 .. n_cand)
  bool have_best = false;   // `best` already covers every document that can hold a key
  if (PRUNE && prune && n_h > 0) {
    const int nh = n_h;
    auto complete = [&](int i) {   // partial -> full score: one load per head term of the query
      float full = acc[i];
      const float* col = hd.head_w + (d0 + i);
      for (int h = 0; h < nh; ++h)
        full = fmaf(h_idf[h], __ldg(col + static_cast<int64_t>(h_slot[h]) * hd.head_ld), full);
      acc[i] = full;
    };
    // Lists the documents with accumulator >= lo_cut (MODE 0) or != 0 (MODE 1) in cand_idx with
    // full warps of survivors (one survivor per warp would cost the whole warp the head loop).
    auto list_docs = [&](int mode, float lo_cut) {
      if (threadIdx.x == 0) n_cand = 0;
      __syncthreads();
      const float4* a4 = reinterpret_cast<const float4*>(acc);
      // (the trip count is the same for every thread: the loop body holds warp collectives)
      for (int b0 = 0; b0 < nd; b0 += kBm25Threads * 4) {
        const int i0 = b0 + threadIdx.x * 4;
        const float4 v = i0 < nd ? a4[i0 >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);   // tile_docs % 4 == 0
        const float e[4] = {v.x, v.y, v.z, v.w};
        unsigned mine = 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool live = i0 + u < nd && (mode == 0 ? e[u] >= lo_cut : e[u] != 0.f);
          mine |= (live ? 1u : 0u) << u;
        }
        if (__any_sync(kFullMask, mine != 0u)) {
          const int cnt = __popc(mine);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
          }
          int base = 0;
          if (lane == 31) base = atomicAdd(&n_cand, incl);
          base = __shfl_sync(kFullMask, base, 31);
          int slot = base + incl - cnt;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if ((mine >> u) & 1u) {
              if (slot < kBm25CandCap) cand_idx[slot] = static_cast<uint16_t>(i0 + u);
              ++slot;
            }
        }
      }
      __syncthreads();
    };
    // theta = the best k-th FULL score any finished tile of this query has published: a lower
    // bound of the query's k-th best (block-uniform; 0 = nothing published yet)
    __shared__ float s_theta;
    if (threadIdx.x == 0) s_theta = theta_g ? __ldcg(theta_g + q) : 0.f;
    __syncthreads();
    const float theta = s_theta;
    const float ub = h_ub * (1.f + 1e-5f);   // slack for the rounding of the completed sum
    if (theta > ub) {
      // a document survives when partial + bound of the head contributions can reach theta
      cut = theta - ub - 1e-6f * theta;
      list_docs(0, cut);
      listed = n_cand <= kBm25CandCap;
      if (listed) {
        for (int c = threadIdx.x; c < n_cand; c += kBm25Threads) complete(cand_idx[c]);
      } else {   // too many survivors to list: complete them in place
        for (int i = threadIdx.x; i < nd; i += kBm25Threads)
          if (acc[i] >= cut) complete(i);
      }
      __syncthreads();
    } else {
      // No useful bound yet.  The documents that contain an essential term are completed first;
      // the k-th best of their full scores is a lower bound of the tile's k-th best.  Only if
      // the head terms alone could still reach it must the other documents be completed too.
      list_docs(1, 0.f);
      bool rest = n_cand > kBm25CandCap;
      if (!rest) {
        for (int c = threadIdx.x; c < n_cand; c += kBm25Threads) {
          const int i = cand_idx[c];
          complete(i);
          const uint64_t key = key_of(i);
          best = key > best ? key : best;
        }
        const uint64_t kth1 = kth_of_thread_bests(best, k, tbest, &s_thr);
        if (kth1 != 0ull && key_score(kth1) > ub) {
          listed = true;   // untouched documents score at most ub: they cannot reach the k-th best
          have_best = true;
          cut = key_score(kth1) - 1e-6f * fabsf(key_score(kth1));   // excludes them from every sweep
          if (!(cut > ub)) { listed = false; have_best = false; cut = -INFINITY; rest = true; }
        } else {
          rest = true;
        }
        if (rest)   // complete everything that has not been completed (accumulator still 0)
          for (int i = threadIdx.x; i < nd; i += kBm25Threads)
            if (acc[i] == 0.f) complete(i);
      } else {
        for (int i = threadIdx.x; i < nd; i += kBm25Threads) complete(i);
      }
      __syncthreads();
    }
  }
  if (!have_best) {
    best = 0ull;
    if (listed) {
      for (int c = threadIdx.x; c < n_cand; c += kBm25Threads) {
        const uint64_t key = key_of(cand_idx[c]);
        best = key > best ? key : best;
      }
    } else {
      for (int i = threadIdx.x; i < nd; i += kBm25Threads) {
        const uint64_t key = key_of(i);
        best = key > best ? key : best;
      }
    }
  }
  if (threadIdx.x == 0) n_sel = 0;
  for (int i = threadIdx.x; i < sel_cap; i += kBm25Threads) sel[i] = 0ull;
  uint64_t thr = 0ull;   // 0: everything that holds a key is collected
  // few listed survivors: rank them all directly, no threshold needed
  const bool small = listed && n_cand <= 64;
  if (!small) thr = kth_of_thread_bests(best, k, tbest, &s_thr);
  else __syncthreads();
  const int n_sweep = listed ? n_cand : nd;
  for (int c = threadIdx.x; c < n_sweep; c += kBm25Threads) {
    const uint64_t key = key_of(listed ? cand_idx[c] : c);
    if (key != 0ull && key >= thr) {
      const int slot = atomicAdd(&n_sel, 1);
      if (slot < sel_cap) sel[slot] = key;   // cannot overflow: <= k threads x ceil(tile / 256) docs
    }
  }
  __syncthreads();
  const int ns = n_sel < sel_cap ? n_sel : sel_cap;
  uint64_t* o = out + out_row * out_stride_q + static_cast<int64_t>(tile) * k;
  if (ns <= kBm25Threads) {
    // rank by counting (keys are unique): one pass, no sort
    for (int i = threadIdx.x; i < k; i += kBm25Threads)
      if (i >= ns) o[i] = 0ull;
    if (threadIdx.x < ns) {
      const uint64_t key = sel[threadIdx.x];
      int rank = 0;
      for (int j = 0; j < ns; ++j) rank += sel[j] > key;
      if (rank < k) o[rank] = key;
      // the tile's k-th best full score bounds the query's k-th best from below: publish it
      if (PRUNE && theta_g && rank == k - 1 && key_score(key) > 0.f)
        atomicMax(reinterpret_cast<int*>(theta_g + q), __float_as_int(key_score(key)));
    }
  } else {
    block_bitonic_sort_desc(sel, next_pow2(ns));
    for (int i = threadIdx.x; i < k; i += blockDim.x) o[i] = i < ns ? sel[i] : 0ull;
    if (PRUNE && theta_g && threadIdx.x == 0 && ns >= k && key_score(sel[k - 1]) > 0.f)
      atomicMax(reinterpret_cast<int*>(theta_g + q), __float_as_int(key_score(sel[k - 1])));
  }
}

template <bool EMIT_ALL, bool PRUNE>
__global__ void __launch_bounds__(kBm25Threads)
bm25_score_kernel(Bm25View ix, Bm25HeadView hd, const int32_t* __restrict__ q_terms,
                  const int32_t* __restrict__ q_offsets, int k,
                  const uint32_t* __restrict__ doc_mask, int tile_docs, int list_cap,
                  uint64_t* __restrict__ out, int64_t out_stride_q, float* __restrict__ theta_g,
                  int tile_stride, int n_sampled, const int32_t* __restrict__ q_list,
                  const int32_t* __restrict__ n_list) {
  if (n_list) {   // device-side list of queries: block row y takes entries y, y + gridDim.y, ...
    pdl_wait();   // (the rerun launch of a chain, launch_bm25_score_listed)
    pdl_trigger();
    const int n = *n_list;
    for (int by = blockIdx.y; by < n; by += gridDim.y) {
      bm25_tile_item<EMIT_ALL, PRUNE>(ix, hd, q_terms, q_offsets, k, doc_mask, tile_docs, list_cap,
                                      out, out_stride_q, theta_g, tile_stride, n_sampled, blockIdx.x,
                                      by, q_list);
      __syncthreads();   // the item's shared memory is no longer read
    }
    return;
  }
  bm25_tile_item<EMIT_ALL, PRUNE>(ix, hd, q_terms, q_offsets, k, doc_mask, tile_docs, list_cap, out,
                                  out_stride_q, theta_g, tile_stride, n_sampled, blockIdx.x,
                                  blockIdx.y, q_list);
}

template <bool EMIT_ALL, bool PRUNE>
static cudaError_t launch_score_t(const Bm25View& ix, const Bm25HeadView& hd, const int32_t* q_terms,
                                  const int32_t* q_offsets, int nq, int k,
                                  const uint32_t* doc_mask, const Bm25Plan& plan, uint64_t* out,
                                  int64_t out_stride_q, float* theta, cudaStream_t stream) {
  if (plan.n_tiles < 1 || nq < 1) return cudaSuccess;
  auto kern = bm25_score_kernel<EMIT_ALL, PRUNE>;
  cudaError_t e =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
  // same L1 / shared-memory split as the dense kernels, so that CTAs of both can share an SM
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  // grid.y is limited to 65535 queries per launch
  for (int q0 = 0; q0 < nq; q0 += 65535) {
    const int nb = nq - q0 < 65535 ? nq - q0 : 65535;
    dim3 grid(plan.n_tiles, nb);
    static const int fold_env = getenv("ANR_BM25_FOLD") ? atoi(getenv("ANR_BM25_FOLD")) : -1;
    const bool fold = fold_env >= 0 ? fold_env != 0 : !plan.beside_dense;
    // sample tiles: every stride-th one (1/16 of a large corpus; a small shard -- 125k documents
    // are 21 tiles -- still gets a sample, or its whole first wave would run without a bound:
    // 0.124 ms per batch of 64 on a 125k-document shard in round 1)
    const int stride = plan.n_tiles >= 32 ? 16 : (plan.n_tiles / 2 > 2 ? plan.n_tiles / 2 : 2);
    constexpr int kMinTilesForSample = 4;
    if (PRUNE && theta && plan.n_tiles >= kMinTilesForSample && !fold) {
      // two launches: sample tiles, then the rest.  Beside the dense pass of a hybrid query this
      // is the better shape: the short first launch leaves the SMs to the dense pass' own short
      // kernels and its main kernel starts early (0.658 vs 0.674 ms per hybrid step); alone, the
      // folded single launch below is ~8 % faster.
      dim3 grid_s((plan.n_tiles + stride - 1) / stride, nb);
      if (plan.phase != 2)
        kern<<<grid_s, kBm25Threads, plan.smem_bytes, stream>>>(
            ix, hd, q_terms, q_offsets + q0, k, doc_mask, plan.tile_docs, plan.list_cap,
            out + q0 * out_stride_q, out_stride_q, theta + q0, stride, -2, nullptr, nullptr);
      if (plan.phase != 1)
        kern<<<grid, kBm25Threads, plan.smem_bytes, stream>>>(
            ix, hd, q_terms, q_offsets + q0, k, doc_mask, plan.tile_docs, plan.list_cap,
            out + q0 * out_stride_q, out_stride_q, theta + q0, stride, -3, nullptr, nullptr);
    } else if (plan.phase == 1) {
      // no separate sample launch in this plan: everything happens in phase 2
    } else if (PRUNE && theta && plan.n_tiles >= kMinTilesForSample && plan.n_tiles <= 65535) {
      // one launch, sample tiles first (see the kernel): no second launch, no idle tail between
      const int n_sampled = (plan.n_tiles + stride - 1) / stride;
      dim3 grid_f(nb, plan.n_tiles);
      kern<<<grid_f, kBm25Threads, plan.smem_bytes, stream>>>(
          ix, hd, q_terms, q_offsets + q0, k, doc_mask, plan.tile_docs, plan.list_cap,
          out + q0 * out_stride_q, out_stride_q, theta + q0, stride, n_sampled, nullptr, nullptr);
    } else {
      kern<<<grid, kBm25Threads, plan.smem_bytes, stream>>>(
          ix, hd, q_terms, q_offsets + q0, k, doc_mask, plan.tile_docs, plan.list_cap,
          out + q0 * out_stride_q, out_stride_q, theta ? theta + q0 : nullptr, 1, -1, nullptr, nullptr);
    }
  }
  return cudaGetLastError();
}

cudaError_t launch_bm25_score_topk(const Bm25View& ix, const Bm25HeadView* hd, const int32_t* q_terms,
                                   const int32_t* q_offsets, int nq, int k,
                                   const uint32_t* doc_mask, const Bm25Plan& plan, uint64_t* cand,
                                   int64_t cand_stride_q, float* theta, cudaStream_t stream) {
  if (k < 1 || k > kMaxFusedK) return cudaErrorInvalidValue;
  if (hd && hd->n_head > 0) {
    if (theta && plan.phase != 2) {
      cudaError_t e = cudaMemsetAsync(theta, 0, static_cast<size_t>(nq) * 4, stream);
      if (e != cudaSuccess) return e;
    }
    return launch_score_t<false, true>(ix, *hd, q_terms, q_offsets, nq, k, doc_mask, plan, cand,
                                       cand_stride_q, theta, stream);
  }
  return launch_score_t<false, false>(ix, Bm25HeadView(), q_terms, q_offsets, nq, k, doc_mask, plan,
                                      cand, cand_stride_q, nullptr, stream);
}

// Exhaustive tiled scan of the queries q_list[0 .. *n_list) (both on the device; at most nq): block
// row b writes the candidates of query q_list[b] to cand[b * cand_stride_q + tile * k + i].  The
// rerun of the queries the candidate-driven path flags (anr_bm25_ms.cu).
cudaError_t launch_bm25_score_listed(const Bm25View& ix, const int32_t* q_terms,
                                     const int32_t* q_offsets, int nq, int k,
                                     const uint32_t* doc_mask, const Bm25Plan& plan, uint64_t* cand,
                                     int64_t cand_stride_q, const int32_t* q_list,
                                     const int32_t* n_list, cudaStream_t stream) {
  if (k < 1 || k > kMaxFusedK) return cudaErrorInvalidValue;
  if (plan.n_tiles < 1 || nq < 1) return cudaSuccess;
  auto kern = bm25_score_kernel<false, false>;
  cudaError_t e =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
  if (e == cudaSuccess)   // (the dense kernels' L1 / shared-memory split: CTAs of both may share an SM)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  // few block rows: the list is usually empty and every CTA returns at once (16 rows of 82 tiles
  // took ~5 us to schedule and exit at the end of the BM25 chain of every step)
  return launch_chain_on(kPdlBm25, kern, dim3(plan.n_tiles, nq < 4 ? nq : 4), dim3(kBm25Threads),
                      static_cast<size_t>(plan.smem_bytes), stream, ix, Bm25HeadView(), q_terms,
                      q_offsets, k, doc_mask, plan.tile_docs, plan.list_cap, cand, cand_stride_q,
                      static_cast<float*>(nullptr), 1, -1, q_list, n_list);
}

cudaError_t launch_bm25_score_all(const Bm25View& ix, const int32_t* q_terms,
                                  const int32_t* q_offsets, int nq, const uint32_t* doc_mask,
                                  const Bm25Plan& plan, uint64_t* keys, int64_t keys_stride_q,
                                  cudaStream_t stream) {
  return launch_score_t<true, false>(ix, Bm25HeadView(), q_terms, q_offsets, nq, 1, doc_mask, plan,
                                     keys, keys_stride_q, nullptr, stream);
}

// ---- dense rows of the head terms ------------------------------------------------------------
constexpr int kBm25HeadPad = 1024;
int64_t bm25_head_ld(int n_docs) {
  return (static_cast<int64_t>(n_docs) + kBm25HeadPad - 1) / kBm25HeadPad * kBm25HeadPad;
}

__global__ void __launch_bounds__(256)
bm25_head_fill_kernel(Bm25View ix, const int32_t* __restrict__ head_terms, float* __restrict__ head_w,
                      float* __restrict__ head_max, int64_t head_ld) {
  const int slot = blockIdx.y;
  const int term = head_terms[slot];
  const int64_t lo = ix.term_ptr[term], hi = ix.term_ptr[term + 1];
  float* row = head_w + static_cast<int64_t>(slot) * head_ld;
  float mx = 0.f;
  for (int64_t p = lo + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < hi;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float w = ix.post_w[p];
    row[ix.post_doc[p]] = w;
    mx = fmaxf(mx, w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
  // weights are positive: their bit patterns order like the values
  if ((threadIdx.x & 31) == 0 && mx > 0.f)
    atomicMax(reinterpret_cast<int*>(head_max + slot), __float_as_int(mx));
}

cudaError_t launch_bm25_head_fill(const Bm25View& ix, const int32_t* head_terms, int n_head,
                                  float* head_w, float* head_max, int64_t head_ld,
                                  cudaStream_t stream) {
  if (n_head < 1) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(head_w, 0, static_cast<size_t>(n_head) * head_ld * 4, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(head_max, 0, static_cast<size_t>(n_head) * 4, stream);
  if (e != cudaSuccess) return e;
  dim3 grid(148 * 2, n_head);
  bm25_head_fill_kernel<<<grid, 256, 0, stream>>>(ix, head_terms, head_w, head_max, head_ld);
  return cudaGetLastError();
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
