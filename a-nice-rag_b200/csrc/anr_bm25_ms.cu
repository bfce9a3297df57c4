// BM25 top-k, candidate-driven (MaxScore with a required set): the default top-k path.
//
// Replaces bm25.get_scores(query_tokens) + the top-k at src/search_engine.py:219, :236-241 (and
// the filtered branch :221-234 through doc_mask) for k <= 128, like the tiled scan in anr_bm25.cu,
// with the same results -- but its cost follows the postings that can matter, not
// (documents x queries): the tiled scan zeroes, fills and sweeps one accumulator per (document,
// query) whatever the pruning saves (ncu, 1M docs x 64 queries: 138 M warp instructions, ~50 %
// issue utilisation, 0.25 GB of DRAM traffic: instruction-bound set-up work).
//
// Every query term t has an upper bound ub_t = idf_t * (largest posting weight of t) on what it
// can add to one document (term_maxw, computed once per index).  Per query:
//   1. PLAN      terms ranked by document frequency; the shortest lists, up to kMsSample postings
//                in all (the last one cut to a prefix), form the sample S1.
//   2. STAGE 1   every S1 posting is a candidate document whose FULL score is computed (below);
//                theta = the k-th best of those scores: a lower bound of the query's k-th best
//                score, because the candidates are distinct documents.
//   3. REQUIRED  terms are set aside, longest list first, while the sum of their bounds stays
//                below theta: a document that holds none of the remaining ("required") terms
//                scores below theta and cannot enter the top-k.
//   4. STAGE 2   every posting of every required list is a candidate: one compare against
//                theta with the bounds of all other terms rejects most; the others are scored in
//                full, abandoning as soon as score-so-far + bounds of the terms still to come
//                fall below theta.  A document found in an EARLIER required list (query order)
//                belongs to that list's candidate: no document is listed twice.
//   5. The survivors (score >= theta) of a query are ranked by the final top-k kernel.
// Full score of a candidate = sum over the query's terms IN QUERY ORDER (duplicates repeat, as in
// the reference loop) of idf_t * w_t(doc), fp32 fma; w_t(doc) is one load from the dense head rows
// for head terms, a binary search in the term's posting list otherwise.  Scores do not depend on
// theta or on timing: results are bit-reproducible.
// Queries the path cannot finish exactly -- fewer than k documents with a positive score (the
// reference then returns zero-score documents too), more survivors than the buffer holds, more
// than kMsMaxTerms terms -- are FLAGGED and rerun through the exhaustive tiled scan on the device
// (anr_api.cu), so the call stays asynchronous and capturable.
// Needs every idf >= 0 (BM25Okapi's epsilon floor guarantees it unless the average idf itself is
// negative; such an index takes the tiled scan).
#include <algorithm>
#include <cstdlib>

#include "anr_internal.h"
#include "anr_topk.cuh"

namespace anr {

constexpr int kMsThreads = 256;
constexpr int kMsThetaSel = 2048;   // theta kernel: sample keys at or above the per-thread-best bound (<= 128 x 16)
constexpr int kMsWarpItems = 64;    // stage 2: consecutive postings a warp's lanes share out at a time

struct __align__(8) MsTerm {
  int64_t lo;           // first posting of the list
  const int32_t* bkt;   // bucket table of the list (MsIndexView), or null
  int32_t len;          // postings (0: the term adds nothing -- unknown id, zero idf, empty list)
  float idf;
  float ub;             // idf * largest weight of the list
  int32_t slot;         // dense head row, or -1
  int32_t s1;           // postings of the list's prefix that are stage-1 candidates
  int32_t s2;           // postings streamed in stage 2: len if required, else 0
  int32_t shift;        // bucket of a document = doc >> shift
  int32_t rank;         // position of this term in the evaluation order (largest bound first)
  float sub;            // sum of ub over the REQUIRED lists ordered before this one (shorter first)
  int32_t pad;
  // evaluation order, indexed by RANK r (not by query position):
  int32_t ord;          // query position of the term evaluated r-th
  float esuf;           // sum of ub over the terms evaluated r-th and later
};
struct __align__(8) MsQuery {
  int64_t s2_total;
  int32_t n;          // query terms
  float theta;
  int32_t s1_total;
  int32_t n_surv;     // survivors appended by stage 2
  int32_t flag;       // 1: rerun through the exhaustive scan
  float tot;          // sum of every term's bound
};

size_t bm25_ms_scratch_bytes(int nq) {
  auto pad = [](size_t b) { return (b + 255) / 256 * 256; };
  return pad(static_cast<size_t>(nq) * sizeof(MsQuery)) +
         pad(static_cast<size_t>(nq) * kMsMaxTerms * sizeof(MsTerm)) +
         pad(static_cast<size_t>(nq) * kMsSample * 8) + pad((static_cast<size_t>(nq) + 1) * 8) + 256 + 1024;
}

// ---- once per index / reweighting: largest posting weight of every term -------------------------
__global__ void __launch_bounds__(256)
bm25_term_max_kernel(const int64_t* __restrict__ term_ptr, const float* __restrict__ post_w,
                     int n_terms, float* __restrict__ term_maxw) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = warp; t < n_terms; t += n_warps) {
    const int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
    float mx = 0.f;
    for (int64_t p = lo + lane; p < hi; p += 32) mx = fmaxf(mx, __ldg(post_w + p));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
    if (lane == 0) term_maxw[t] = mx;
  }
}
// *neg = 1 when some idf is negative
__global__ void __launch_bounds__(256)
bm25_neg_idf_kernel(const float* __restrict__ idf, int n_terms, int32_t* __restrict__ neg) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += gridDim.x * blockDim.x)
    if (idf[t] < 0.f) *neg = 1;
}
cudaError_t launch_bm25_term_max(const Bm25View& ix, float* term_maxw, int32_t* neg_idf,
                                 cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(neg_idf, 0, 4, stream);
  if (e != cudaSuccess || ix.n_terms < 1) return e;
  const int blocks = static_cast<int>(std::min<int64_t>((static_cast<int64_t>(ix.n_terms) + 7) / 8, 148 * 32));
  bm25_term_max_kernel<<<blocks, 256, 0, stream>>>(ix.term_ptr, ix.post_w, ix.n_terms, term_maxw);
  bm25_neg_idf_kernel<<<std::min((ix.n_terms + 255) / 256, 592), 256, 0, stream>>>(ix.idf, ix.n_terms,
                                                                                  neg_idf);
  return cudaGetLastError();
}

// ---- once per index: bucket tables ------------------------------------------------------------
// "Is document d in the list of term t, and with what weight?" is the inner operation of the whole
// path.  A binary search costs log2(len) DEPENDENT loads (18 for 250k postings, each an L2 round
// trip); a list of kMsBucketMin+ postings therefore gets a table of the positions where the
// document ranges [b << shift, (b + 1) << shift) start, with 2-4 postings per range on average:
// the lookup is one load of two neighbouring entries + one or two steps inside one sector + the
// weight.  (1/4 to 1/2 of an int32 per posting.)  Measured (profiles/r2_call13/14): what bounds the
// path is the NUMBER of divergent loads (one sector per lane and load), not their latency --
// resolving all terms of a candidate side by side (8 lookups in three rounds of independent loads)
// was slower (stage 2: 180 against 122 us) than resolving them one by one and stopping early.
__global__ void __launch_bounds__(128)
bm25_bucket_fill_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ post_doc,
                        const int64_t* __restrict__ bkt_off, const uint8_t* __restrict__ bkt_shift,
                        int n_docs, int32_t* __restrict__ bkt) {
  const int t = blockIdx.x;
  const int shift = bkt_shift[t];
  if (shift == 0xff) return;
  const int64_t lo = term_ptr[t];
  const int len = static_cast<int>(term_ptr[t + 1] - lo);
  int32_t* table = bkt + bkt_off[t];
  const int n_buckets = ((n_docs - 1) >> shift) + 1;   // entries 0 .. n_buckets (the last = len)
  const int32_t* docs = post_doc + lo;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const int b = docs[i] >> shift;
    const int b_prev = i > 0 ? docs[i - 1] >> shift : -1;
    for (int bb = b_prev + 1; bb <= b; ++bb) table[bb] = i;   // first posting of every range up to mine
  }
  const int b_last = len > 0 ? docs[len - 1] >> shift : -1;
  for (int bb = b_last + 1 + threadIdx.x; bb <= n_buckets; bb += blockDim.x) table[bb] = len;
}
cudaError_t launch_bm25_bucket_fill(const Bm25View& ix, const int64_t* bkt_off, const uint8_t* bkt_shift,
                                    int32_t* bkt, cudaStream_t stream) {
  if (ix.n_terms < 1) return cudaSuccess;
  bm25_bucket_fill_kernel<<<ix.n_terms, 128, 0, stream>>>(ix.term_ptr, ix.post_doc, bkt_off, bkt_shift,
                                                         ix.n_docs, bkt);
  return cudaGetLastError();
}
// host side of the layout: shift of a list of `len` postings (0xff: no table)
int bm25_bucket_shift(int64_t len, int n_docs) {
  if (len < kMsBucketMin || n_docs < 2) return 0xff;
  int s = 0;
  while ((static_cast<int64_t>(n_docs) >> s) > len / 2) ++s;   // len / 4 < ranges <= len / 2
  return s;
}
int64_t bm25_bucket_entries(int shift, int n_docs) {
  return shift == 0xff ? 0 : ((static_cast<int64_t>(n_docs) - 1) >> shift) + 2;
}

// ---- 1. plan: one warp per query --------------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(kMsThreads)
ms_plan_kernel(Bm25View ix, Bm25HeadView hd, MsIndexView mx, const int32_t* __restrict__ q_terms,
               const int32_t* __restrict__ q_offsets, int nq, int sample,
               MsQuery* __restrict__ queries, MsTerm* __restrict__ terms,
               int32_t* __restrict__ ticket) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (kMsThreads / 32) + warp;
  if (blockIdx.x == 0 && threadIdx.x == 0) { ticket[0] = 0; ticket[1] = 0; }
  if (q >= nq) return;
  const int t0 = q_offsets[q], n = q_offsets[q + 1] - t0;
  MsQuery* Q = queries + q;
  if (n > kMsMaxTerms || n <= 0) {
    if (lane == 0) {
      Q->s2_total = 0; Q->n = 0; Q->theta = 0.f; Q->s1_total = 0; Q->n_surv = 0; Q->tot = 0.f;
      Q->flag = n > kMsMaxTerms ? 1 : 0;   // no terms: no result at all (search_engine.py:216-217)
    }
    return;
  }
  int64_t lo[2] = {0, 0};
  const int32_t* bk[2] = {nullptr, nullptr};
  int len[2] = {0, 0}, slot[2] = {-1, -1}, shift[2] = {0, 0};
  float idf[2] = {0.f, 0.f}, ub[2] = {0.f, 0.f};
  bool negative = false;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int j = lane + 32 * e;
    if (j < n) {
      const int term = q_terms[t0 + j];
      if (term >= 0 && term < ix.n_terms) {
        const float v = ix.idf[term];
        negative |= v < 0.f;
        if (v > 0.f) {   // `idf.get(q) or 0`: a zero idf adds nothing
          const int64_t a = ix.term_ptr[term], b = ix.term_ptr[term + 1];
          if (b > a) {
            lo[e] = a;
            len[e] = static_cast<int>(b - a);
            idf[e] = v;
            ub[e] = v * mx.term_maxw[term];
            if (hd.n_head > 0) {
              const int s = hd.slot_of[term];
              slot[e] = s != 0xff ? s : -1;
            }
            const int sh = mx.bkt_shift[term];
            if (sh != 0xff) { bk[e] = mx.bkt + mx.bkt_off[term]; shift[e] = sh; }
          }
        }
      }
    }
  }
  negative = __any_sync(kFullMask, negative);
  // sample order = (list length, position); evaluation order = (bound descending, position)
  int64_t before[2] = {0, 0};
  int rank[2] = {0, 0};
  float esuf[2] = {0.f, 0.f};
  float tot = 0.f;
  for (int jj = 0; jj < n; ++jj) {
    const int owner = jj & 31, e2 = jj >> 5;
    const int len_jj = __shfl_sync(kFullMask, e2 ? len[1] : len[0], owner);
    const float ub_jj = __shfl_sync(kFullMask, e2 ? ub[1] : ub[0], owner);
    tot += ub_jj;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = lane + 32 * e;
      if (len_jj > 0 && len[e] > 0 && (len_jj < len[e] || (len_jj == len[e] && jj < j)))
        before[e] += len_jj;
      const bool jj_first = ub_jj > ub[e] || (ub_jj == ub[e] && jj < j);   // jj is evaluated before j
      if (jj_first) ++rank[e]; else esuf[e] += ub_jj;                      // (j itself lands here)
    }
  }
  int s1_sum = 0;
  MsTerm* T = terms + static_cast<size_t>(q) * kMsMaxTerms;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int j = lane + 32 * e;
    int s1 = 0;
    if (len[e] > 0 && before[e] < sample)
      s1 = static_cast<int>(min(static_cast<int64_t>(len[e]), sample - before[e]));
    s1_sum += s1;
    if (j < n) {
      MsTerm& t = T[j];
      t.lo = lo[e]; t.bkt = bk[e]; t.len = len[e]; t.idf = idf[e]; t.ub = ub[e]; t.slot = slot[e];
      t.s1 = s1; t.s2 = 0; t.shift = shift[e]; t.rank = rank[e];
      T[rank[e]].ord = j;          // (ranks are a permutation of 0 .. n-1)
      T[rank[e]].esuf = esuf[e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s1_sum += __shfl_xor_sync(kFullMask, s1_sum, o);
  if (lane == 0) {
    Q->s2_total = 0; Q->n = n; Q->theta = 0.f; Q->s1_total = negative ? 0 : s1_sum; Q->n_surv = 0;
    Q->flag = negative ? 1 : 0; Q->tot = tot;
  }
}

// ---- position of `doc` in a posting list, or -1 ---------------------------------------------------
// bucket bounds (one round trip) -> the bucket's ids, four neighbours side by side (one round trip;
// a crowded bucket or a list without a table is first narrowed by bisection)
__device__ __forceinline__ int ms_find(const int32_t* __restrict__ docs, int len,
                                       const int32_t* __restrict__ bkt, int shift, int doc) {
  int l = 0, h = len;
  if (bkt) {
    const int b = doc >> shift;
    l = __ldg(bkt + b);
    h = __ldg(bkt + b + 1);
  }
  while (h - l > 4) {
    const int mid = (l + h) >> 1;
    if (__ldg(docs + mid) < doc) l = mid + 1; else h = mid + 1;
  }
  int pos = -1;
#pragma unroll
  for (int v = 0; v < 4; ++v)
    if (l + v < h && __ldg(docs + l + v) == doc) pos = l + v;
  return pos;
}

// weight of `doc` for term t (0: absent; weights are positive): one load from the dense row of a
// head term; else bucket bounds, then the bucket's ids AND their weights side by side -- two round
// trips instead of three (the path is bound by the latency of dependent DRAM accesses)
__device__ __forceinline__ float ms_weight(const Bm25View& ix, const Bm25HeadView& hd, const MsTerm& t,
                                           int doc) {
  if (t.slot >= 0) return __ldg(hd.head_w + static_cast<int64_t>(t.slot) * hd.head_ld + doc);
  const int32_t* docs = ix.post_doc + t.lo;
  const float* ws = ix.post_w + t.lo;
  int l = 0, h = t.len;
  if (t.bkt) {
    const int b = doc >> t.shift;
    l = __ldg(t.bkt + b);
    h = __ldg(t.bkt + b + 1);
  }
  while (h - l > 4) {
    const int mid = (l + h) >> 1;
    if (__ldg(docs + mid) < doc) l = mid + 1; else h = mid + 1;
  }
  float w = 0.f;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const bool in = l + v < h;
    const int d = in ? __ldg(docs + l + v) : -1;
    const float wv = in ? __ldg(ws + l + v) : 0.f;
    if (d == doc) w = wv;
  }
  return w;
}
// the reference's sum: every term of the query in query order (duplicates repeat), fp32 fma
__device__ __forceinline__ float ms_full_score(const Bm25View& ix, const Bm25HeadView& hd,
                                               const MsTerm* __restrict__ T, int n, int j_self,
                                               float c_self, int doc) {
  float full = 0.f;
  for (int jj = 0; jj < n; ++jj) {
    const MsTerm& t = T[jj];
    if (t.len == 0) continue;
    const float w = jj == j_self ? c_self : ms_weight(ix, hd, t, doc);
    if (w > 0.f) full = fmaf(t.idf, w, full);
  }
  return full;
}

__device__ __forceinline__ bool ms_allowed(const uint32_t* __restrict__ mask, int doc) {
  return !mask || ((__ldg(mask + (doc >> 5)) >> (doc & 31)) & 1u);
}

// ---- 2. stage 1: full score of every sample posting ---------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(kMsThreads)
ms_stage1_kernel(Bm25View ix, Bm25HeadView hd, const uint32_t* __restrict__ doc_mask,
                 const MsQuery* __restrict__ queries, const MsTerm* __restrict__ terms,
                 uint64_t* __restrict__ s1keys) {
  pdl_wait();
  pdl_trigger();
  const int q = blockIdx.y;
  const int i0 = blockIdx.x * kMsThreads + threadIdx.x;
  const int n1 = queries[q].s1_total;
  if (i0 >= n1) return;
  const int n = queries[q].n;
  const MsTerm* T = terms + static_cast<size_t>(q) * kMsMaxTerms;
  int j = 0, i = i0;
  while (j < n - 1 && i >= T[j].s1) { i -= T[j].s1; ++j; }
  const int64_t pos = T[j].lo + i;
  const int doc = __ldg(ix.post_doc + pos);
  uint64_t key = 0ull;
  if (ms_allowed(doc_mask, doc)) {
    // a document sampled through an EARLIER list (query order) is that list's candidate
    bool dup = false;
    for (int jj = 0; jj < j && !dup; ++jj) {
      const MsTerm& t = T[jj];
      if (t.s1 > 0) {
        const int p = ms_find(ix.post_doc + t.lo, t.len, t.bkt, t.shift, doc);
        dup = p >= 0 && p < t.s1;
      }
    }
    if (!dup) {
      const float full = ms_full_score(ix, hd, T, n, j, __ldg(ix.post_w + pos), doc);
      if (full > 0.f) key = make_key(full, static_cast<uint32_t>(doc));
    }
  }
  s1keys[static_cast<size_t>(q) * kMsSample + i0] = key;
}

// ---- 3. theta + required set: one CTA per query; the last CTA to finish lays the queries' -------
//         stage-2 postings end to end (q_base = exclusive prefix of s2_total)
template <int VARIANT>
__global__ void __launch_bounds__(kMsThreads)
ms_theta_kernel(int k, int nq, MsQuery* __restrict__ queries, MsTerm* __restrict__ terms,
                const uint64_t* __restrict__ s1keys, int64_t* __restrict__ q_base,
                int32_t* __restrict__ ticket) {
  __shared__ uint64_t tbest[kMsThreads];
  __shared__ uint64_t sel[kMsThetaSel];
  __shared__ uint64_t s_kth;
  __shared__ int64_t part[kMsThreads];
  __shared__ int n_sel;
  __shared__ bool last;
  // no early trigger: stage 2 is next, and its 5 CTAs per SM hold every register of an SM -- made
  // resident while this kernel runs they took SMs ahead of the dense main kernel of the step, whose
  // CTAs then waited for them (the dense kernel's span in a step: 0.32 -> 0.35 ms)
  pdl_wait();
  const int q = blockIdx.x, lane = threadIdx.x & 31;
  MsQuery* Q = queries + q;
  const int n1 = Q->s1_total, n = Q->n;
  uint64_t best = 0ull;
  for (int i = threadIdx.x; i < n1; i += kMsThreads) {
    const uint64_t v = s1keys[static_cast<size_t>(q) * kMsSample + i];
    best = v > best ? v : best;
  }
  uint64_t kth = block_kth_of_thread_bests<kMsThreads / 32>(best, k, tbest, &s_kth);
  // kth bounds the sample's k-th best from below; the exact k-th best (a higher theta: fewer
  // required lists, earlier rejections) is among the <= k * ceil(n1 / 256) keys at or above it
  if (kth != 0ull && n1 > kMsThreads) {
    if (threadIdx.x == 0) n_sel = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n1; i += kMsThreads) {
      const uint64_t v = s1keys[static_cast<size_t>(q) * kMsSample + i];
      if (v >= kth) {
        const int slot = atomicAdd(&n_sel, 1);
        if (slot < kMsThetaSel) sel[slot] = v;
      }
    }
    __syncthreads();
    const int m = min(n_sel, kMsThetaSel);
    if (n_sel <= kMsThetaSel && m >= k) {   // (always, by the bound above; else keep the lower bound)
      const int p2 = next_pow2(m);
      for (int i = m + threadIdx.x; i < p2; i += kMsThreads) sel[i] = 0ull;
      block_bitonic_sort_desc(sel, p2);
      kth = sel[k - 1];
    }
  }
  const float theta = kth != 0ull ? key_score(kth) : 0.f;
  if (threadIdx.x < 32) {
    MsTerm* T = terms + static_cast<size_t>(q) * kMsMaxTerms;
    int len[2] = {0, 0};
    float ub[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = lane + 32 * e;
      if (j < n) { len[e] = T[j].len; ub[e] = T[j].ub; }
    }
    // set aside, longest list first: cum = bounds of the lists ordered at or before mine
    float cum[2] = {0.f, 0.f};
    for (int jj = 0; jj < n; ++jj) {
      const int owner = jj & 31, e2 = jj >> 5;
      const int len_jj = __shfl_sync(kFullMask, e2 ? len[1] : len[0], owner);
      const float ub_jj = __shfl_sync(kFullMask, e2 ? ub[1] : ub[0], owner);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = lane + 32 * e;
        if (len_jj > 0 && len[e] > 0 && (len_jj > len[e] || (len_jj == len[e] && jj <= j)))
          cum[e] += ub_jj;
      }
    }
    const float tot = Q->tot;
    const bool flagged = Q->flag != 0;
    int64_t s2_sum = 0;
    int s2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool aside = cum[e] + 2e-5f * tot < theta;
      s2[e] = (len[e] > 0 && !aside && !flagged) ? len[e] : 0;
      s2_sum += s2[e];
    }
    // A document belongs to the SHORTEST required list that holds it.  sub = bounds of the required
    // lists ordered before mine: none of them can add to a candidate that is mine to score.
    float sub[2] = {0.f, 0.f};
    for (int jj = 0; jj < n; ++jj) {
      const int owner = jj & 31, e2 = jj >> 5;
      const int s2_jj = __shfl_sync(kFullMask, e2 ? s2[1] : s2[0], owner);
      const float ub_jj = __shfl_sync(kFullMask, e2 ? ub[1] : ub[0], owner);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = lane + 32 * e;
        if (s2_jj > 0 && s2[e] > 0 && (s2_jj < s2[e] || (s2_jj == s2[e] && jj < j))) sub[e] += ub_jj;
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = lane + 32 * e;
      if (j < n) { T[j].s2 = s2[e]; T[j].sub = sub[e]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2_sum += __shfl_xor_sync(kFullMask, s2_sum, o);
    if (lane == 0) {
      Q->theta = theta;
      Q->s2_total = s2_sum;
      __threadfence();
      last = atomicAdd(ticket, 1) == nq - 1;
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // exclusive prefix over the queries (volatile: written by other CTAs of this launch)
  const int per = (nq + kMsThreads - 1) / kMsThreads;
  const int a = min(nq, static_cast<int>(threadIdx.x) * per), b = min(nq, a + per);
  int64_t s = 0;
  for (int i = a; i < b; ++i) s += *reinterpret_cast<volatile int64_t*>(&queries[i].s2_total);
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < kMsThreads; o <<= 1) {
    const int64_t v = static_cast<int>(threadIdx.x) >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int64_t run = part[threadIdx.x] - s;
  for (int i = a; i < b; ++i) {
    q_base[i] = run;
    run += *reinterpret_cast<volatile int64_t*>(&queries[i].s2_total);
  }
  if (threadIdx.x == kMsThreads - 1) q_base[nq] = part[kMsThreads - 1];
}

// ---- 4. stage 2: every posting of the required lists --------------------------------------------
// One candidate per LANE, and a lane that is done with its candidate takes the next posting at
// once: every trip of the loop is one lookup per lane, whatever stage each lane's candidate is in
// (bounds / ownership).  With lock-step rounds of 32 candidates a warp waited for its
// longest-lived candidate in every round (ncu: 33 % of the warp slots active, 115 us).
template <int VARIANT>
__global__ void __launch_bounds__(kMsThreads, 5)
ms_stage2_kernel(Bm25View ix, Bm25HeadView hd, const uint32_t* __restrict__ doc_mask, int nq, int cap,
                 MsQuery* __restrict__ queries, const MsTerm* __restrict__ terms,
                 const int64_t* __restrict__ q_base, uint64_t* __restrict__ surv) {
  pdl_wait();   // (no early trigger: the ranking kernel need not park beside this long one)
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int64_t total = q_base[nq];
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * (kMsThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (kMsThreads / 32);
  int64_t chunk = gw * kMsWarpItems;                       // warp-uniform: the chunk being handed out
  int64_t cursor = chunk, cend = min(total, chunk + kMsWarpItems);
  // per-lane candidate
  bool have = false;
  int q = -1, j = 0, n = 0, doc = 0, phase = 0, idx = 0, self_s2 = 0;
  int64_t q_lo = 0, q_hi = 0;                              // items of query q: [q_lo, q_hi)
  int64_t seg_lo = 0, seg_hi = 0, self_lo = 0;             // items of its list j: [seg_lo, seg_hi)
  float self_idf = 0.f, self_rem = 0.f;
  const MsTerm* T = terms;
  float c_self = 0.f, partial = 0.f, remaining = 0.f, theta = 0.f, slack = 0.f, tot = 0.f;
  float wv[kMsMaxTerms];   // weights of my candidate's terms as they are looked up (local memory)

  auto before = [&](const MsTerm& t, int jj) {   // required list ordered before mine: it owns shared documents
    return t.s2 > 0 && (t.s2 < self_s2 || (t.s2 == self_s2 && jj < j));
  };

  for (;;) {
    // ---- lanes without a candidate take the next postings (up to four attempts while half the warp
    //      is free: most postings die on the first compare, before any lookup)
#pragma unroll 1
    for (int attempt = 0; attempt < 4; ++attempt) {
      const unsigned want = __ballot_sync(kFullMask, !have);
      if (want == 0u || (attempt >= 2 && __popc(want) < 16)) break;
      if (cursor >= cend) {   // next chunk of this warp
        chunk += n_warps * kMsWarpItems;
        cursor = chunk;
        cend = min(total, chunk + kMsWarpItems);
      }
      if (cursor >= total) break;
      const int avail = static_cast<int>(cend - cursor);
      const int mine_rank = __popc(want & lt_mask);
      const bool take = !have && mine_rank < avail;
      const int64_t item = cursor + mine_rank;
      cursor += min(__popc(want), avail);
      if (take) {
        if (item >= seg_hi || item < seg_lo) {   // (usually the list my last posting came from)
          if (item >= q_hi || item < q_lo) {     // the query that holds the item: last q with q_base[q] <= item
            int l = 0, h = nq;
            while (h - l > 1) {
              const int mid = (l + h) >> 1;
              if (q_base[mid] <= item) l = mid; else h = mid;
            }
            q = l;
            q_lo = q_base[q];
            q_hi = q_base[q + 1];
            T = terms + static_cast<size_t>(q) * kMsMaxTerms;
            n = queries[q].n;
            theta = queries[q].theta;
            tot = queries[q].tot;
            slack = 2e-5f * tot;
          }
          int64_t r = item - q_lo;
          j = 0;
          while (r >= T[j].s2) { r -= T[j].s2; ++j; }
          seg_lo = item - r;
          seg_hi = seg_lo + T[j].s2;
          self_s2 = T[j].s2;
          self_lo = T[j].lo;
          self_idf = T[j].idf;
          self_rem = tot - T[j].ub - T[j].sub;
        }
        const int64_t p_self = self_lo + (item - seg_lo);
        doc = __ldg(ix.post_doc + p_self);
        c_self = __ldg(ix.post_w + p_self);
        // What a document that is MINE to score can reach at most: its own posting + the bounds
        // of the other terms, the required lists ordered before mine excluded (a document found
        // there is that list's candidate).
        partial = self_idf * c_self;
        remaining = self_rem;
        if (ms_allowed(doc_mask, doc) && !(partial + remaining + slack < theta)) {
          have = true;
          phase = 0;
          idx = 0;
        }
      }
    }
    if (__ballot_sync(kFullMask, have) == 0u) {
      if (cursor >= total || (cursor >= cend && chunk + n_warps * kMsWarpItems >= total)) break;
      continue;
    }
    // ---- which term does my candidate resolve next?
    int look = -1;
    if (have) {
      if (phase == 0) {   // the terms that can add, largest bound first
        while (idx < n) {
          const int jj = T[idx].ord;
          const MsTerm& t = T[jj];
          if (t.len == 0) { idx = n; break; }   // (terms that add nothing sort last)
          if (jj != j && !before(t, jj)) { look = jj; break; }
          ++idx;
        }
        if (look < 0) { phase = 1; idx = 0; }
      }
      if (phase == 1) {   // does a required list ordered before mine hold the document?
        while (idx < n && !(idx != j && before(T[idx], idx))) ++idx;
        if (idx < n) {
          look = idx;
        } else {
          // every term is resolved (the weights of the terms that can add were kept as they were
          // looked up; a list ordered before mine does not hold the document): the reference's
          // sum -- query order, duplicates repeat, fp32 fma -- and a survivor if it reaches theta
          float full = 0.f;
          for (int jj = 0; jj < n; ++jj) {
            const MsTerm& t = T[jj];
            if (t.len == 0) continue;
            const float w = jj == j ? c_self : (before(t, jj) ? 0.f : wv[jj]);
            if (w > 0.f) full = fmaf(t.idf, w, full);
          }
          if (full >= theta && full > 0.f) {
            const int s = atomicAdd(&queries[q].n_surv, 1);
            if (s < cap) surv[static_cast<size_t>(q) * cap + s] = make_key(full, static_cast<uint32_t>(doc));
            else queries[q].flag = 1;
          }
          have = false;
        }
      }
    }
    // ---- one lookup per lane
    if (look >= 0) {
      const MsTerm& t = T[look];
      float w = 0.f;
      bool found = false;
      if (phase != 1) w = ms_weight(ix, hd, t, doc);
      else found = ms_find(ix.post_doc + t.lo, t.len, t.bkt, t.shift, doc) >= 0;
      if (phase == 0) {
        wv[look] = w;
        remaining -= t.ub;
        if (w > 0.f) partial = fmaf(t.idf, w, partial);
        if (partial + fmaxf(remaining, 0.f) + slack < theta) have = false;
      } else {
        if (found) have = false;   // that list owns the document
      }
      ++idx;
    }
  }
}

// ---- which queries go through the exhaustive scan: flags -> ascending list (one whole CTA) ------
__device__ __forceinline__ void ms_collect_flags(const MsQuery* __restrict__ queries,
                                                 const int32_t* __restrict__ q_offsets, int nq, int k,
                                                 int cap, int32_t* __restrict__ n_flagged,
                                                 int32_t* __restrict__ flagged) {
  __shared__ int count;
  __shared__ int warp_off[kMsThreads / 32];
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int base = 0; base < nq; base += kMsThreads) {
    const int q = base + threadIdx.x;
    bool f = false;
    if (q < nq) {
      const int n_surv = *reinterpret_cast<const volatile int32_t*>(&queries[q].n_surv);
      const int flag = *reinterpret_cast<const volatile int32_t*>(&queries[q].flag);
      const bool empty = q_offsets[q + 1] == q_offsets[q];
      // fewer than k documents with a positive score: the reference ranks zero-score documents too
      f = !empty && (flag != 0 || n_surv < k || n_surv > cap);
    }
    const unsigned ballot = __ballot_sync(kFullMask, f);
    if (lane == 0) warp_off[warp] = __popc(ballot);
    __syncthreads();
    int off = count;
    for (int w = 0; w < warp; ++w) off += warp_off[w];
    if (f) flagged[off + __popc(ballot & ((1u << lane) - 1))] = q;
    __syncthreads();
    if (threadIdx.x == 0)
      for (int w = 0; w < kMsThreads / 32; ++w) count += warp_off[w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_flagged = count;
}

// ---- 5. the survivors of a query, ranked: one CTA per query --------------------------------------
// Usually a few dozen keys: one key per thread, rank by counting (keys are unique).  More than 256
// (a weak theta: up to kMsSurvivors): the k-th largest of the per-thread bests bounds the k-th best
// from below, the keys at or above it -- at most k * ceil(ns / 256) <= 2048 -- are collected and
// sorted.  (Ranking 4000 survivors by counting took 0.6 ms per launch at 256 queries.)
constexpr int kMsFinalSel = 2048;
template <int VARIANT>
__global__ void __launch_bounds__(kMsThreads)
ms_final_kernel(const MsQuery* __restrict__ queries, const uint64_t* __restrict__ surv, int cap, int k,
                int nq, TopkOut o, int32_t* __restrict__ ticket, int32_t* __restrict__ n_flagged,
                int32_t* __restrict__ flagged) {
  __shared__ uint64_t keys[kMsThreads];
  __shared__ uint64_t sel[kMsFinalSel];
  __shared__ uint64_t s_kth;
  __shared__ int n_sel;
  __shared__ bool last;
  pdl_wait();
  pdl_trigger();
  const int q = blockIdx.x;
  const int ns = min(queries[q].n_surv, cap);
  const uint64_t* sv = surv + static_cast<size_t>(q) * cap;
  const bool dead = o.q_offsets && o.q_offsets[q + 1] == o.q_offsets[q];   // a query without terms
  if (ns <= kMsThreads) {
    const uint64_t key = static_cast<int>(threadIdx.x) < ns ? sv[threadIdx.x] : 0ull;
    keys[threadIdx.x] = key;
    __syncthreads();
    int rank = 0;
    for (int i = 0; i < ns; ++i) rank += keys[i] > key;
    for (int i = ns + threadIdx.x; i < k; i += kMsThreads) emit_entry(0ull, q * o.stride_q + i, o);
    if (key != 0ull && rank < k) emit_entry(dead ? 0ull : key, q * o.stride_q + rank, o);
  } else {
    uint64_t best = 0ull;
    for (int i = threadIdx.x; i < ns; i += kMsThreads) {
      const uint64_t v = sv[i];
      best = v > best ? v : best;
    }
    if (threadIdx.x == 0) n_sel = 0;
    const uint64_t thr = block_kth_of_thread_bests<kMsThreads / 32>(best, k, keys, &s_kth);
    for (int i = threadIdx.x; i < ns; i += kMsThreads) {
      const uint64_t v = sv[i];
      if (v >= thr) {
        const int slot = atomicAdd(&n_sel, 1);
        if (slot < kMsFinalSel) sel[slot] = v;   // cannot overflow: <= k threads x ceil(ns / 256) keys
      }
    }
    __syncthreads();
    const int m = min(n_sel, kMsFinalSel);
    const int p2 = next_pow2(m < 2 ? 2 : m);
    for (int i = m + threadIdx.x; i < p2; i += kMsThreads) sel[i] = 0ull;
    block_bitonic_sort_desc(sel, p2);
    for (int i = threadIdx.x; i < k; i += kMsThreads)
      emit_entry((i < m && !dead) ? sel[i] : 0ull, q * o.stride_q + i, o);
  }
  if (threadIdx.x == 0) {
    if (o.counts) o.counts[q * o.count_stride] = dead ? 0 : min(ns, k);
    last = atomicAdd(ticket, 1) == nq - 1;
  }
  __syncthreads();
  // the last CTA lists the queries for the exhaustive scan (stage 2 is complete: kernel boundary)
  if (last) ms_collect_flags(queries, o.q_offsets, nq, k, cap, n_flagged, flagged);
}

// plan | stage 1 | theta + required lists | stage 2 | ranking + flagged list, one instantiation
template <int VARIANT, int EARLY>
static void ms_launch_chain(const DeviceProps& dp, const Bm25View& ix, const Bm25HeadView& hd,
                            const MsIndexView& mx, const int32_t* q_terms, const int32_t* q_offsets,
                            int nq, int k, const uint32_t* doc_mask, int sample, MsQuery* queries,
                            MsTerm* terms, uint64_t* s1keys, int64_t* q_base, int32_t* ticket,
                            uint64_t* surv, const TopkOut& out, int32_t* n_flagged, int32_t* flagged,
                            cudaStream_t stream, const cudaEvent_t* marks) {
  auto mark = [&](int i) { if (marks && marks[i]) cudaEventRecord(marks[i], stream); };
  const int wpb = kMsThreads / 32;
  launch_chain_on(kPdlBm25, ms_plan_kernel<EARLY>, dim3((nq + wpb - 1) / wpb), dim3(kMsThreads), 0, stream, ix, hd,
               mx, q_terms, q_offsets, nq, sample, queries, terms, ticket);
  mark(0);
  launch_chain_on(kPdlBm25, ms_stage1_kernel<EARLY>, dim3(sample / kMsThreads, nq), dim3(kMsThreads), 0, stream, ix,
               hd, doc_mask, queries, terms, s1keys);
  mark(1);
  launch_chain_on(kPdlBm25, ms_theta_kernel<EARLY>, dim3(nq), dim3(kMsThreads), 0, stream, k, nq, queries, terms,
               s1keys, q_base, ticket);
  mark(2);
  launch_chain_on(kPdlBm25, ms_stage2_kernel<VARIANT>, dim3(dp.sm_count * 5), dim3(kMsThreads), 0, stream, ix, hd,
               doc_mask, nq, static_cast<int>(kMsSurvivors), queries, terms, q_base, surv);
  mark(3);
  launch_chain_on(kPdlBm25, ms_final_kernel<VARIANT>, dim3(nq), dim3(kMsThreads), 0, stream, queries, surv,
               static_cast<int>(kMsSurvivors), k, nq, out, ticket + 1, n_flagged, flagged);
}

cudaError_t launch_bm25_maxscore(const DeviceProps& dp, const Bm25View& ix, const Bm25HeadView& hd,
                                 const MsIndexView& mx, const int32_t* q_terms,
                                 const int32_t* q_offsets, int nq, int k, const uint32_t* doc_mask,
                                 unsigned char* scratch, uint64_t* surv, const TopkOut& out,
                                 int32_t* n_flagged, int32_t* flagged, cudaStream_t stream,
                                 bool beside_dense, const cudaEvent_t* marks) {
  if (nq < 1) return cudaSuccess;
  auto pad = [](size_t b) { return (b + 255) / 256 * 256; };
  MsQuery* queries = reinterpret_cast<MsQuery*>(scratch);
  scratch += pad(static_cast<size_t>(nq) * sizeof(MsQuery));
  MsTerm* terms = reinterpret_cast<MsTerm*>(scratch);
  scratch += pad(static_cast<size_t>(nq) * kMsMaxTerms * sizeof(MsTerm));
  uint64_t* s1keys = reinterpret_cast<uint64_t*>(scratch);
  scratch += pad(static_cast<size_t>(nq) * kMsSample * 8);
  int64_t* q_base = reinterpret_cast<int64_t*>(scratch);
  scratch += pad((static_cast<size_t>(nq) + 1) * 8);
  int32_t* ticket = reinterpret_cast<int32_t*>(scratch);
  // Two instantiations of every kernel.  VARIANT 1 runs beside the dense pass of a hybrid step and
  // asks for the dense kernels' L1 / shared-memory split (all shared): CTAs that ask for different
  // splits cannot share an SM -- without it the stage kernels took the SMs first and the dense CTAs
  // waited for them to drain (0.487 against 0.460 ms per step, profiles/r2_call24/25).  VARIANT 0
  // runs alone with the default split: the lookups live on L1 hits (72 %), and the all-shared
  // split costs them 40 % (0.133 -> 0.190 ms per batch of 64).
  if (beside_dense) {
    static const bool set = [] {
      auto prefer = [](const void* f) {
        cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      };
      prefer(reinterpret_cast<const void*>(ms_plan_kernel<1>));
      prefer(reinterpret_cast<const void*>(ms_stage1_kernel<1>));
      prefer(reinterpret_cast<const void*>(ms_theta_kernel<1>));
      prefer(reinterpret_cast<const void*>(ms_stage2_kernel<1>));
      prefer(reinterpret_cast<const void*>(ms_final_kernel<1>));
      return true;
    }();
    (void)set;
  }
  // stage-1 candidates per query (ANR_MS_SAMPLE, a profiling knob; at most kMsSample)
  static const int sample_env = getenv("ANR_MS_SAMPLE") ? atoi(getenv("ANR_MS_SAMPLE")) : kMsSample;
  const int sample = std::min(std::max(sample_env, kMsThreads), kMsSample) / kMsThreads * kMsThreads;
  static const bool early_default = getenv("ANR_MS_EARLY_SPLIT") && atoi(getenv("ANR_MS_EARLY_SPLIT")) == 0;
  if (beside_dense && early_default)
    ms_launch_chain<1, 0>(dp, ix, hd, mx, q_terms, q_offsets, nq, k, doc_mask, sample, queries, terms, s1keys,
                          q_base, ticket, surv, out, n_flagged, flagged, stream, marks);
  else if (beside_dense)
    ms_launch_chain<1, 1>(dp, ix, hd, mx, q_terms, q_offsets, nq, k, doc_mask, sample, queries, terms, s1keys,
                       q_base, ticket, surv, out, n_flagged, flagged, stream, marks);
  else
    ms_launch_chain<0, 0>(dp, ix, hd, mx, q_terms, q_offsets, nq, k, doc_mask, sample, queries, terms, s1keys,
                       q_base, ticket, surv, out, n_flagged, flagged, stream, marks);
  return cudaGetLastError();
}

}  // namespace anr
