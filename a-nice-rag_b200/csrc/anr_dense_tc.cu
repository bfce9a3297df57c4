// Dense inner-product scan for LARGE query batches on the 5th-generation tensor cores.
//
// At more than a few queries per pass the scan stops being a memory-bound matrix-vector
// product (anr_dense.cu) and becomes a dense contraction  S[rows, queries] = E · Qᵀ.  This
// kernel computes it with tcgen05.mma.kind::tf32 directly on the fp32 corpus -- the tensor
// core reads the fp32 words as tf32 (10-bit mantissa), so no shadow copy of the corpus is
// needed and HBM traffic stays at rows x D x 4 bytes per pass of 32 queries -- and never
// writes the score matrix: the epilogue reads the accumulator tile from TMEM, rejects almost
// every score with one compare against a running per-(warp, query) threshold and keeps a short
// candidate list per query.  tf32 scores are only used to NOMINATE candidates: the rescoring
// kernel below recomputes every candidate within a rigorous error margin of the k-th best in
// exact fp32 and ranks on those scores, and flags a query for the exact scan when a candidate
// list could have dropped a row inside the margin (so results equal the exact path's).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer: 2-D tensor-map loads (SWIZZLE_128B) of [128 rows x 32 floats] corpus
//             boxes into a ring; the 32 x D query block is loaded once and stays resident
//   warp 1    allocates TMEM, then one elected lane issues 4 x (M=128, N=32, K=8) MMAs per box
//             and commits them to the ring's "empty" barriers / the accumulator "full" barrier
//   warps 2-5 epilogue: tcgen05.ld of the finished accumulator (lane = corpus row, column =
//             query), candidate selection; two accumulators so tile t+1 runs under it
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "anr_internal.h"
#include "anr_tc.cuh"
#include "anr_topk.cuh"

namespace anr {

constexpr int kTcRows = 128;        // corpus rows per tile = UMMA M
constexpr int kTcQueries = 32;      // queries per pass = UMMA N
constexpr int kTcSlab = 32;         // floats per K slab = one 128-byte swizzle row
constexpr int kTcThreads = 192;     // producer warp, MMA warp, 4 epilogue warps
constexpr int kTcEpiWarps = 4;
constexpr int kTcMaxStages = 8;
constexpr int kTcABytes = kTcRows * kTcSlab * 4;      // 16 KB per stage
constexpr int kTcBSlabBytes = kTcQueries * kTcSlab * 4;  // 4 KB per query slab

// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both
// K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) |
                              (static_cast<uint32_t>(kTcQueries >> 3) << 17) |
                              (static_cast<uint32_t>(kTcRows >> 4) << 24);

struct TcLayout {
  int b_off, ring_off, list_off, bar_off, total_bytes;
  int n_stages, n_slabs, kl;
};


// ---- the scan --------------------------------------------------------------------------
// EMIT = false: cand is [kTcQueries][grid][kTcEpiWarps][kl] candidate keys with tf32 scores
//               (0 = empty slot), selected above the starting thresholds thr0 (nullable)
// EMIT = true : cand is [kTcQueries][n] keys of EVERY row (0 = masked), used on a small sample
template <bool EMIT>
__global__ void __launch_bounds__(kTcThreads, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                int64_t n, const uint32_t* __restrict__ mask, const uint64_t* __restrict__ thr0,
                uint64_t* __restrict__ cand, TcLayout L) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* b_smem = smem + L.b_off;       // n_slabs x [32 queries x 128 B], swizzled
  unsigned char* ring = smem + L.ring_off;      // n_stages x [128 rows x 128 B], swizzled
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + L.list_off);  // [4 warps][32 queries][kl]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty = full + kTcMaxStages;
  uint64_t* b_full = empty + kTcMaxStages;
  uint64_t* acc_full = b_full + 1;              // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (n + kTcRows - 1) / kTcRows;
  const int64_t my_tiles =
      n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  for (int i = threadIdx.x; i < kTcEpiWarps * kTcQueries * L.kl; i += blockDim.x) lists[i] = 0ull;
  if (threadIdx.x == 0) {
    for (int s = 0; s < L.n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kTcEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // 64 TMEM columns: two 128 x 32 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer ----
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(L.n_slabs) * kTcBSlabBytes);
      for (int kb = 0; kb < L.n_slabs; ++kb)
        tma_load_2d(b_smem + static_cast<size_t>(kb) * kTcBSlabBytes, &map_b, kb * kTcSlab, 0,
                    b_full);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int row0 = static_cast<int>((blockIdx.x + it * gridDim.x) * kTcRows);
        for (int kb = 0; kb < L.n_slabs; ++kb) {
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], kTcABytes);
          tma_load_2d(ring + static_cast<size_t>(s) * kTcABytes, &map_a, kb * kTcSlab, row0,
                      &full[s]);
          if (++s == L.n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    if (lane == 0) {
      mbar_wait(b_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int a = static_cast<int>(it & 1);
        const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
        mbar_wait(&acc_empty[a], aph ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(a * kTcQueries);
        for (int kb = 0; kb < L.n_slabs; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t da = tc_smem_desc(smem_u32(ring + static_cast<size_t>(s) * kTcABytes));
          const uint64_t db = tc_smem_desc(smem_u32(b_smem + static_cast<size_t>(kb) * kTcBSlabBytes));
#pragma unroll
          for (int kk = 0; kk < kTcSlab / 8; ++kk)  // K = 8 tf32 = 32 bytes per MMA
            tc_mma_tf32(tmem_d, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2),
                        kTcIdesc, (kb | kk) != 0 ? 1u : 0u);
          tc_commit(&empty[s]);  // stage reusable once these MMAs have read it
          if (++s == L.n_stages) { s = 0; ph ^= 1u; }
        }
        tc_commit(&acc_full[a]);  // accumulator complete
      }
    }
  } else {
    // ---- epilogue: TMEM lane = corpus row, column = query ----
    const int ew = warp - 2;               // list owner index
    const int quad = warp & 3;             // TMEM lane quadrant this warp may read
    uint64_t* my_lists = lists + static_cast<size_t>(ew) * kTcQueries * L.kl;
    // starting thresholds from the sample pre-pass (a lower bound of every query's global
    // kl-th best): without them each warp would insert ~kl * ln(rows / kl) rows per query
    uint64_t thr[kTcQueries];
#pragma unroll
    for (int j = 0; j < kTcQueries; ++j) thr[j] = thr0 ? thr0[j] : 0ull;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int a = static_cast<int>(it & 1);
      const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      uint32_t v[32];
      tc_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                      static_cast<uint32_t>(a * kTcQueries),
                  v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);

      const int64_t row = (blockIdx.x + it * gridDim.x) * kTcRows + quad * 32 + lane;
      bool ok = row < n;
      if (ok && mask) ok = (__ldg(mask + (row >> 5)) >> (row & 31)) & 1u;
#pragma unroll
      for (int j = 0; j < kTcQueries; ++j) {
        const uint64_t key = ok ? make_key(__uint_as_float(v[j]), static_cast<uint32_t>(row)) : 0ull;
        if (EMIT) {
          if (row < n) cand[static_cast<int64_t>(j) * n + row] = key;
          continue;
        }
        unsigned pending = __ballot_sync(kFullMask, key > thr[j]);
        while (pending) {  // rare
          const int src = __ffs(pending) - 1;
          pending &= pending - 1;
          const uint64_t c = __shfl_sync(kFullMask, key, src);
          if (c > thr[j]) {
            // the list minimum is 0 until the list is full: never let it lower the bar
            const uint64_t nm = warp_list_insert_cold(my_lists + j * L.kl, L.kl, c, lane);
            thr[j] = nm > thr[j] ? nm : thr[j];
          }
        }
      }
    }
    __syncwarp();
    // lists -> global: cand[q][cta][warp][kl]
    if (!EMIT)
      for (int i = lane; i < kTcQueries * L.kl; i += 32) {
        const int q = i / L.kl, e = i % L.kl;
        cand[((static_cast<int64_t>(q) * gridDim.x + blockIdx.x) * kTcEpiWarps + ew) * L.kl + e] =
            my_lists[i];
      }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64)
                 : "memory");
  }
}

// ---- 64 queries per pass: a CTA pair shares every corpus tile --------------------------------
// Cluster of 2 CTAs.  Rank r keeps queries [32r, 32r + 32) resident and runs its own MMAs,
// accumulators and epilogue exactly like dense_tc_kernel, but each corpus box is fetched from
// HBM ONCE per pair: rank r loads rows [64r, 64r + 64) of the box and the TMA multicasts them
// into BOTH CTAs' rings, so HBM traffic per query halves.  A stage may be refilled only after
// both CTAs' MMAs have read it: every commit arrives on the "empty" barrier of both CTAs.

// map_a: box [64 rows x 32 floats]; map_b: [64 queries, ld], box [32 x 32].
// cand: [64 queries][n_clusters][kTcEpiWarps][kl]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
dense_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a,
                     const __grid_constant__ CUtensorMap map_b, int64_t n,
                     const uint32_t* __restrict__ mask, const uint64_t* __restrict__ thr0,
                     uint64_t* __restrict__ cand, TcLayout L) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* b_smem = smem + L.b_off;
  unsigned char* ring = smem + L.ring_off;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + L.list_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty = full + kTcMaxStages;
  uint64_t* b_full = empty + kTcMaxStages;
  uint64_t* acc_full = b_full + 1;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int64_t n_tiles = (n + kTcRows - 1) / kTcRows;
  const int64_t my_tiles =
      n_tiles > cluster ? (n_tiles - cluster + n_clusters - 1) / n_clusters : 0;

  for (int i = threadIdx.x; i < kTcEpiWarps * kTcQueries * L.kl; i += blockDim.x) lists[i] = 0ull;
  if (threadIdx.x == 0) {
    for (int s = 0; s < L.n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);   // one commit from each CTA of the pair
    }
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kTcEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(L.n_slabs) * kTcBSlabBytes);
      for (int kb = 0; kb < L.n_slabs; ++kb)
        tma_load_2d(b_smem + static_cast<size_t>(kb) * kTcBSlabBytes, &map_b, kb * kTcSlab,
                    rank * kTcQueries, b_full);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int row0 = static_cast<int>((cluster + it * n_clusters) * kTcRows);
        for (int kb = 0; kb < L.n_slabs; ++kb) {
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], kTcABytes);   // own half + the peer's half
          tma_load_2d_mcast(ring + static_cast<size_t>(s) * kTcABytes + rank * (kTcABytes / 2),
                            &map_a, kb * kTcSlab, row0 + rank * (kTcRows / 2), &full[s], 0x3);
          if (++s == L.n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(b_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int a = static_cast<int>(it & 1);
        const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
        mbar_wait(&acc_empty[a], aph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(a * kTcQueries);
        for (int kb = 0; kb < L.n_slabs; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t da = tc_smem_desc(smem_u32(ring + static_cast<size_t>(s) * kTcABytes));
          const uint64_t db = tc_smem_desc(smem_u32(b_smem + static_cast<size_t>(kb) * kTcBSlabBytes));
#pragma unroll
          for (int kk = 0; kk < kTcSlab / 8; ++kk)
            tc_mma_tf32(tmem_d, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2),
                        kTcIdesc, (kb | kk) != 0 ? 1u : 0u);
          tc_commit_mcast(&empty[s], 0x3);
          if (++s == L.n_stages) { s = 0; ph ^= 1u; }
        }
        tc_commit(&acc_full[a]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;
    uint64_t* my_lists = lists + static_cast<size_t>(ew) * kTcQueries * L.kl;
    uint64_t thr[kTcQueries];
#pragma unroll
    for (int j = 0; j < kTcQueries; ++j) thr[j] = thr0 ? thr0[rank * kTcQueries + j] : 0ull;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int a = static_cast<int>(it & 1);
      const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      uint32_t v[32];
      tc_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                      static_cast<uint32_t>(a * kTcQueries),
                  v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);

      const int64_t row = (cluster + it * n_clusters) * kTcRows + quad * 32 + lane;
      bool ok = row < n;
      if (ok && mask) ok = (__ldg(mask + (row >> 5)) >> (row & 31)) & 1u;
#pragma unroll
      for (int j = 0; j < kTcQueries; ++j) {
        const uint64_t key = ok ? make_key(__uint_as_float(v[j]), static_cast<uint32_t>(row)) : 0ull;
        unsigned pending = __ballot_sync(kFullMask, key > thr[j]);
        while (pending) {
          const int src = __ffs(pending) - 1;
          pending &= pending - 1;
          const uint64_t c = __shfl_sync(kFullMask, key, src);
          if (c > thr[j]) {
            const uint64_t nm = warp_list_insert_cold(my_lists + j * L.kl, L.kl, c, lane);
            thr[j] = nm > thr[j] ? nm : thr[j];
          }
        }
      }
    }
    __syncwarp();
    for (int i = lane; i < kTcQueries * L.kl; i += 32) {
      const int q = rank * kTcQueries + i / L.kl, e = i % L.kl;
      cand[((static_cast<int64_t>(q) * n_clusters + cluster) * kTcEpiWarps + ew) * L.kl + e] =
          my_lists[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64)
                 : "memory");
  }
}

// ---- exact rescoring -----------------------------------------------------------------
// One CTA per query.  Among the m = grid * 4 * kl nominated candidates: T = k-th best tf32
// score; every candidate with tf32 score >= T - 2*eps is rescored in exact fp32 (eps bounds
// |tf32 score - exact score|), the exact top-k is emitted.  flags[q] = 1 when exactness cannot be
// guaranteed from the lists (a full list whose minimum lies inside the margin may have dropped
// a relevant row, or more than kTcRescoreCap candidates fall inside the margin): the caller
// then reruns that query through the exact scan.
constexpr int kTcRescoreCap = 2048;
constexpr int kTcRescoreThreads = 1024;

// Sample pre-pass: thr0[q] = key that rejects every score <= the kl-th best tf32 score among the
// sample's candidates.  The kl-th best of a subset never exceeds the kl-th best of the whole
// corpus, so rows at or below it cannot be among the global top-kl.
__global__ void __launch_bounds__(kTcRescoreThreads)
dense_tc_thr_kernel(const uint64_t* __restrict__ cand, int m, int kl,
                    uint64_t* __restrict__ thr0) {
  // (kl = the rank asked for, <= kTcRescoreThreads)
  // kl-th largest of the per-thread bests: kl distinct rows reach it, so it is a lower
  // bound of the sample's (hence the corpus') kl-th best -- all a starting threshold needs
  __shared__ uint64_t best[kTcRescoreThreads];
  const int q = blockIdx.x;
  const uint64_t* c = cand + static_cast<int64_t>(q) * m;
  uint64_t b = 0ull;
  for (int i = threadIdx.x; i < m; i += kTcRescoreThreads) {
    const uint64_t v = c[i];
    b = v > b ? v : b;
  }
  best[threadIdx.x] = b;
  block_bitonic_sort_desc(best, kTcRescoreThreads);
  if (threadIdx.x == 0) {
    const uint64_t kth = best[kl - 1];
    thr0[q] = kth ? (kth | 0xffffffffull) : 0ull;
  }
}

__global__ void __launch_bounds__(kTcRescoreThreads)
dense_tc_rescore_kernel(const uint64_t* __restrict__ cand, int n_lists, int kl,
                        const float* __restrict__ emb, int ld, const float* __restrict__ q_dev,
                        int k, float eps_scale, const uint64_t* __restrict__ thr0, TopkOut o,
                        int32_t* __restrict__ flags, const int32_t* __restrict__ cnt, int cap,
                        int32_t* __restrict__ ticket, int32_t* __restrict__ n_flagged,
                        int32_t* __restrict__ flagged) {
  __shared__ uint64_t top[kTcRescoreThreads];    // per-thread best tf32 keys
  __shared__ uint64_t sel[kTcRescoreCap];        // candidates inside the margin, then exact keys
  __shared__ int n_sel, bad;
  __shared__ float q_norm2;
  pdl_wait();
  pdl_trigger();
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = kTcRescoreThreads / 32;
  // cnt == nullptr: n_lists bounded candidate lists of kl keys (0 = empty slot);
  // cnt != nullptr: one append buffer of cap keys per query holding min(cnt[q], cap) keys
  const uint64_t* c = cand + static_cast<int64_t>(q) * (cnt ? cap : n_lists * kl);
  const int m = cnt ? min(cnt[q], cap) : n_lists * kl;
  if (cnt) n_lists = 0;
  const float* qv = q_dev + static_cast<size_t>(q) * ld;

  if (threadIdx.x == 0) { n_sel = 0; bad = (cnt && cnt[q] > cap) ? 1 : 0; q_norm2 = 0.f; }
  __syncthreads();
  // |q|^2
  float part = 0.f;
  for (int i = threadIdx.x; i < ld; i += blockDim.x) part = fmaf(qv[i], qv[i], part);
  part = warp_sum(part);
  if (lane == 0) atomicAdd(&q_norm2, part);
  // T = k-th largest of the per-thread bests: a lower bound of the k-th best tf32 score
  // (k distinct candidates reach it).  A lower T only widens the rescored margin, so exactness
  // is unaffected and no list of candidates has to be maintained here.
  __shared__ uint64_t s_kth;
  uint64_t kth;
  {
    uint64_t b = 0ull;
#pragma unroll 8
    for (int i = threadIdx.x; i < m; i += kTcRescoreThreads) {   // (independent loads: one round trip)
      const uint64_t v = __ldg(c + i);
      b = v > b ? v : b;
    }
    kth = block_kth_of_thread_bests<kTcRescoreThreads / 32>(b, k, top, &s_kth);
  }
  // error bound of a tf32 product sum: each operand loses < 2^-10 relative (13 mantissa bits
  // dropped), accumulation noise D * 2^-22 -> |err| <= 2.5e-3 * |q| * |e| (Cauchy-Schwarz)
  const float eps = eps_scale * sqrtf(q_norm2);
  // fewer than k candidates exist at all: everything nominated is rescored
  const float cut = kth ? key_score(kth) - 2.f * eps : -INFINITY;
  __syncthreads();
  // rows rejected by the pre-pass threshold are safe only if that threshold lies below the margin
  if (threadIdx.x == 0 && thr0 && thr0[q] != 0ull && key_score(thr0[q]) >= cut) bad = 1;
  for (int l = threadIdx.x; l < n_lists; l += blockDim.x) {   // could a list have dropped a row?
    uint64_t mn = ~0ull;
    for (int e = 0; e < kl; ++e) {
      const uint64_t v = c[static_cast<int64_t>(l) * kl + e];
      mn = v < mn ? v : mn;
    }
    if (mn != 0ull && key_score(mn) >= cut) bad = 1;
  }
#pragma unroll 8
  for (int i = threadIdx.x; i < m; i += kTcRescoreThreads) {
    const uint64_t v = __ldg(c + i);
    if (v != 0ull && key_score(v) >= cut) {
      const int slot = atomicAdd(&n_sel, 1);
      if (slot < kTcRescoreCap) sel[slot] = v;
    }
  }
  __syncthreads();
  int ns = n_sel;
  if (ns > kTcRescoreCap) {
    if (threadIdx.x == 0) bad = 1;
    ns = kTcRescoreCap;
  }
  // exact fp32 inner products, one warp per candidate, four candidates in flight per warp (the
  // rows are random 4 KB reads: latency, not bytes).  The summation order per row is fixed.
  for (int i = warp; i < ns; i += 4 * n_warps) {
    uint32_t row[4];
    const float* e[4];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int iu = i + u * n_warps;
      row[u] = key_id(sel[iu < ns ? iu : i]);
      e[u] = emb + static_cast<size_t>(row[u]) * ld;
    }
#pragma unroll 2
    for (int col = lane * 4; col < ld; col += 128) {
      const float4 qq = __ldg(reinterpret_cast<const float4*>(qv + col));
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(e[u] + col));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u] = fmaf(v[u].x, qq.x, acc[u]);
        acc[u] = fmaf(v[u].y, qq.y, acc[u]);
        acc[u] = fmaf(v[u].z, qq.z, acc[u]);
        acc[u] = fmaf(v[u].w, qq.w, acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = warp_sum(acc[u]);
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int iu = i + u * n_warps;
        if (iu < ns) sel[iu] = make_key(acc[u], row[u]);
      }
    }
  }
  __syncthreads();
  uint64_t key = 0ull;
  int my_rank = threadIdx.x;
  if (ns <= kTcRescoreThreads) {
    // rank by counting (keys carry unique row ids): no sort, no barriers
    my_rank = k;   // = nothing to emit
    if (threadIdx.x < ns) {
      key = sel[threadIdx.x];
      int rank = 0;
      for (int j = 0; j < ns; ++j) rank += sel[j] > key;
      my_rank = rank;
    }
    if (my_rank >= k) key = 0ull;
    // slots past the number of candidates are empty
    if (threadIdx.x >= ns && threadIdx.x < k) my_rank = threadIdx.x;
  } else {
    const int sp2 = next_pow2(ns);
    for (int i = ns + threadIdx.x; i < sp2; i += blockDim.x) sel[i] = 0ull;
    block_bitonic_sort_desc(sel, sp2);
    if (threadIdx.x < k) key = sel[threadIdx.x];
  }
  if (my_rank < k) {
    const int64_t slot = q * o.stride_q + my_rank;
    const bool valid = key != 0ull;
    const uint32_t id = key_id(key);
    const uint32_t out_id =
        valid ? (o.id_map ? static_cast<uint32_t>(o.id_map[id]) : static_cast<uint32_t>(id + o.id_base))
              : 0u;
    if (o.keys) o.keys[slot] = valid ? ((key & 0xffffffff00000000ull) | (0xffffffffu - out_id)) : 0ull;
    if (o.scores) o.scores[slot] = valid ? key_score(key) : 0.f;
    if (o.ids) o.ids[slot] = valid ? static_cast<int32_t>(out_id) : -1;
  }
  const int n_out = __syncthreads_count(key != 0ull);
  if (threadIdx.x == 0) {
    if (o.counts) o.counts[q * o.count_stride] = n_out;
    flags[q] = bad;
  }
  // ticket != nullptr (one launch covers the whole batch, <= 1024 queries): the last CTA to finish
  // lists the flagged queries in ascending order -- what compact_flags_kernel does in a launch of
  // its own (~4 us of every step, for a list that is almost always empty)
  if (ticket) {
    __shared__ bool s_last;
    __shared__ int s_off[kTcRescoreThreads / 32];
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(ticket, 1) == static_cast<int>(gridDim.x) - 1;
    }
    __syncthreads();
    if (s_last) {
      const int t = threadIdx.x;
      const bool f = t < static_cast<int>(gridDim.x) &&
                     *reinterpret_cast<const volatile int32_t*>(flags + t) != 0;
      const unsigned ballot = __ballot_sync(kFullMask, f);
      if (lane == 0) s_off[warp] = __popc(ballot);
      __syncthreads();
      int off = 0, total = 0;
      for (int w = 0; w < kTcRescoreThreads / 32; ++w) {
        if (w < warp) off += s_off[w];
        total += s_off[w];
      }
      if (f) flagged[off + __popc(ballot & ((1u << lane) - 1u))] = t;
      if (t == 0) *n_flagged = total;
    }
  }
}

// max |row| over the corpus (for the error bound); result in *out (must be zeroed first)
__global__ void __launch_bounds__(256)
row_norm_max_kernel(const float* __restrict__ emb, int64_t n, int ld, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best = 0.f;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < n;
       row += static_cast<int64_t>(gridDim.x) * 8) {
    const float* e = emb + row * ld;
    float acc = 0.f;
    for (int col = lane * 4; col < ld; col += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(e + col));
      acc = fmaf(a.x, a.x, acc);
      acc = fmaf(a.y, a.y, acc);
      acc = fmaf(a.z, a.z, acc);
      acc = fmaf(a.w, a.w, acc);
    }
    acc = warp_sum(acc);
    best = fmaxf(best, acc);
  }
  // non-negative floats order like their bit patterns
  if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best)));
}

cudaError_t launch_row_norm_max(const float* emb, int64_t n, int ld, float* out,
                                cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(out, 0, 4, stream);
  if (e != cudaSuccess || n <= 0) return e;
  int64_t blocks = (n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  row_norm_max_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(emb, n, ld, out);
  return cudaGetLastError();
}

// ---- host side ---------------------------------------------------------------------------
static int64_t tc_sample_rows(const DeviceProps& dp, int64_t n, int k);
// Rows of the sample pre-pass: the planned sample, shrunk to a quarter of a small corpus;
// 0 = corpus too small for a pre-pass to pay off.
static int64_t tc_prepass_rows(const DeviceProps& dp, int64_t n, int k) {
  int64_t rows = tc_sample_rows(dp, n, k);
  if (n < 4 * rows) rows = n / 4 / kTcRows * kTcRows;
  return rows >= 8 * kTcRows ? rows : 0;
}

static inline int tc_align_up(int v, int a) { return (v + a - 1) / a * a; }

// Per-(warp, query) candidate list length.  With the pre-pass threshold only a handful of rows
// per list survive (see tc_sample_rows), so the lists stay short whatever k is; a list that does
// fill up inside the margin flags the query for the exact scan.
static int tc_list_len(int k) { return k <= 10 ? 16 : 32; }

// Rank of the sample's score used as starting threshold, and the sample size that leaves
// about rank * n / sample <= 4096 rows per query above it (~7 per list).
static int tc_thr_rank(int k) { return k <= 10 ? 16 : k + 32; }
static int64_t tc_sample_rows(const DeviceProps& dp, int64_t n, int k) {
  const int64_t wave = static_cast<int64_t>(dp.sm_count) * kTcRows;
  int64_t want = (static_cast<int64_t>(tc_thr_rank(k)) * n + 4095) / 4096;
  if (want < wave) want = wave;
  return (want + wave - 1) / wave * wave;
}

static bool make_tc_layout(const DeviceProps& dp, int ld, int k, TcLayout* L) {
  if (ld % kTcSlab != 0 || k < 1 || k > 128) return false;
  L->n_slabs = ld / kTcSlab;
  L->kl = tc_list_len(k);
  L->b_off = 0;
  L->ring_off = tc_align_up(L->n_slabs * kTcBSlabBytes, 1024);
  const int list_bytes = kTcEpiWarps * kTcQueries * L->kl * 8;
  const int bar_bytes = (2 * kTcMaxStages + 1 + 4) * 8 + 16;
  const int avail = dp.max_smem_optin - 1024 /* alignment slack */ - L->ring_off - list_bytes -
                    bar_bytes - 256;
  int n_stages = avail / kTcABytes;
  if (n_stages < 3) return false;
  if (n_stages > kTcMaxStages) n_stages = kTcMaxStages;
  L->n_stages = n_stages;
  L->list_off = L->ring_off + n_stages * kTcABytes;
  L->bar_off = tc_align_up(L->list_off + list_bytes, 16);
  L->total_bytes = L->bar_off + bar_bytes;
  return true;
}

bool dense_tc_supported(const DeviceProps& dp, int64_t n, int ld, int k) {
  if (getenv("ANR_DISABLE_TC")) return false;
  TcLayout L;
  return n >= kTcRows && n < (1ll << 31) && make_tc_layout(dp, ld, k, &L);
}

int dense_tc_queries_per_pass() { return kTcQueries; }

// scratch one pass needs: candidates of the main pass + of the sample pre-pass + thresholds
size_t dense_tc_cand_keys(const DeviceProps& dp, int64_t n, int k) {
  // sized for the 64-query pair pass (the 32-query pass needs half of it)
  return 2 * (static_cast<size_t>(kTcQueries) * dp.sm_count * kTcEpiWarps * tc_list_len(k) +
              static_cast<size_t>(kTcQueries) * tc_sample_rows(dp, n, k) + kTcQueries);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) ==
            cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, ld] (row-major), box [box_rows, 32 floats], 128-byte swizzle.
static bool encode_map(CUtensorMap* map, const float* base, int64_t rows, int ld, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  const cuuint32_t box[2] = {kTcSlab, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// One pass: the 32 queries at q_dev ([32, ld], zero rows for padding) against the whole corpus.
// Writes candidates to cand (dense_tc_cand_keys keys) and the exact top-k of the first
// n_real queries through `out` (already offset to the first query), flags[0..n_real).
cudaError_t launch_dense_tc(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                            const float* q_dev, int n_real, int k, const uint32_t* mask,
                            float emb_norm_max, uint64_t* cand, const TopkOut& out, int32_t* flags,
                            cudaEvent_t ev_start, cudaEvent_t ev_stop, cudaStream_t stream) {
  TcLayout L;
  if (!make_tc_layout(dp, ld, k, &L)) return cudaErrorInvalidConfiguration;
  CUtensorMap map_a, map_b;
  if (!encode_map(&map_a, emb, n, ld, kTcRows) || !encode_map(&map_b, q_dev, kTcQueries, ld, kTcQueries))
    return cudaErrorInvalidValue;
  const int smem = L.total_bytes + 1024;  // room to align the dynamic base to 1024 bytes
  cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(dense_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem);
  if (e != cudaSuccess) return e;
  const int64_t n_tiles = (n + kTcRows - 1) / kTcRows;
  const int grid = static_cast<int>(n_tiles < dp.sm_count ? n_tiles : dp.sm_count);
  const size_t pass_keys = static_cast<size_t>(kTcQueries) * dp.sm_count * kTcEpiWarps * L.kl;
  uint64_t* cand_sample = cand + pass_keys;
  uint64_t* thr0 = nullptr;
  // sample pre-pass over the first tile of every CTA (worth it from ~8 tiles per CTA on)
  const int64_t n_sample = tc_prepass_rows(dp, n, k);
  if (n_sample > 0) {
    thr0 = cand_sample + static_cast<size_t>(kTcQueries) * n_sample;
    const int grid_s = static_cast<int>(std::min<int64_t>(dp.sm_count, n_sample / kTcRows));
    dense_tc_kernel<true><<<grid_s, kTcThreads, smem, stream>>>(map_a, map_b, n_sample, mask,
                                                                nullptr, cand_sample, L);
    dense_tc_thr_kernel<<<kTcQueries, kTcRescoreThreads, 0, stream>>>(
        cand_sample, static_cast<int>(n_sample), tc_thr_rank(k), thr0);
  }
  if (ev_start) cudaEventRecord(ev_start, stream);   // brackets the main scan kernel only
  dense_tc_kernel<false><<<grid, kTcThreads, smem, stream>>>(map_a, map_b, n, mask, thr0, cand, L);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dense_tc_rescore_kernel<<<n_real, kTcRescoreThreads, 0, stream>>>(
      cand, grid * kTcEpiWarps, L.kl, emb, ld, q_dev, k, 2.5e-3f * emb_norm_max, thr0, out, flags,
      nullptr, 0, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}

// One pass over the corpus for 64 queries at q_dev ([64, ld]) with the CTA-pair kernel.
cudaError_t launch_dense_tc_pair(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                 const float* q_dev, int n_real, int k, const uint32_t* mask,
                                 float emb_norm_max, uint64_t* cand, const TopkOut& out,
                                 int32_t* flags, cudaEvent_t ev_start, cudaEvent_t ev_stop,
                                 cudaStream_t stream) {
  TcLayout L;
  if (!make_tc_layout(dp, ld, k, &L)) return cudaErrorInvalidConfiguration;
  CUtensorMap map_a_half, map_a, map_b64, map_b0, map_b1;
  if (!encode_map(&map_a_half, emb, n, ld, kTcRows / 2) || !encode_map(&map_a, emb, n, ld, kTcRows) ||
      !encode_map(&map_b64, q_dev, 2 * kTcQueries, ld, kTcQueries) ||
      !encode_map(&map_b0, q_dev, kTcQueries, ld, kTcQueries) ||
      !encode_map(&map_b1, q_dev + static_cast<size_t>(kTcQueries) * ld, kTcQueries, ld, kTcQueries))
    return cudaErrorInvalidValue;
  const int smem = L.total_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(dense_tc_pair_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(dense_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem);
  if (e != cudaSuccess) return e;
  const int n_clusters = dp.sm_count / 2;
  const size_t pass_keys = 2 * static_cast<size_t>(kTcQueries) * n_clusters * kTcEpiWarps * L.kl;
  uint64_t* cand_sample = cand + pass_keys;
  uint64_t* thr0 = nullptr;
  const int64_t n_sample = tc_prepass_rows(dp, n, k);
  if (n_sample > 0) {   // two 32-query sample pre-passes (emit mode), one threshold kernel
    thr0 = cand_sample + 2 * static_cast<size_t>(kTcQueries) * n_sample;
    const int grid_s = static_cast<int>(std::min<int64_t>(dp.sm_count, n_sample / kTcRows));
    dense_tc_kernel<true><<<grid_s, kTcThreads, smem, stream>>>(map_a, map_b0, n_sample, mask,
                                                                nullptr, cand_sample, L);
    dense_tc_kernel<true><<<grid_s, kTcThreads, smem, stream>>>(
        map_a, map_b1, n_sample, mask, nullptr,
        cand_sample + static_cast<size_t>(kTcQueries) * n_sample, L);
    dense_tc_thr_kernel<<<2 * kTcQueries, kTcRescoreThreads, 0, stream>>>(
        cand_sample, static_cast<int>(n_sample), tc_thr_rank(k), thr0);
  }
  if (ev_start) cudaEventRecord(ev_start, stream);
  dense_tc_pair_kernel<<<2 * n_clusters, kTcThreads, smem, stream>>>(map_a_half, map_b64, n, mask,
                                                                     thr0, cand, L);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dense_tc_rescore_kernel<<<n_real, kTcRescoreThreads, 0, stream>>>(
      cand, n_clusters * kTcEpiWarps, L.kl, emb, ld, q_dev, k, 2.5e-3f * emb_norm_max, thr0, out,
      flags, nullptr, 0, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}

// Rescoring over per-query append buffers (the GEMM path, anr_dense_gemm.cu).
cudaError_t launch_dense_tc_rescore_append(const uint64_t* cand, const int32_t* cnt, int cap,
                                           const float* emb, int ld, const float* q_dev, int n_real,
                                           int k, float eps_scale, const uint64_t* thr_key,
                                           const TopkOut& out, int32_t* flags, cudaStream_t stream,
                                           int32_t* ticket, int32_t* n_flagged, int32_t* flagged) {
  if (n_real > kTcRescoreThreads) ticket = nullptr;   // (the last CTA lists one flag per thread)
  return launch_chain(dense_tc_rescore_kernel, dim3(n_real), dim3(kTcRescoreThreads), 0, stream, cand,
                      0, 0, emb, ld, q_dev, k, eps_scale, thr_key, out, flags, cnt, cap, ticket,
                      n_flagged, flagged);
}

bool dense_tc_pair_enabled() { return getenv("ANR_DISABLE_TC_PAIR") == nullptr; }

}  // namespace anr
