// Dense inner-product search for query batches of more than 32 queries as a TILED GEMM on the
// 5th-generation tensor cores:  S[rows, queries] = E · Qᵀ  with 128-row x NQ-query output tiles
// (NQ = 64, 128 or 256), both operands streamed per 128-byte K slab by TMA into one shared-memory
// ring, accumulators in TMEM (double buffered: 2 x NQ columns), tcgen05.mma issued by one thread.
//
// Differences from the 32-query scan in anr_dense_tc.cu (which keeps its query block resident):
//   * the query slab travels with the corpus slab (it comes out of L2: the whole query block is
//     a few MB), so the tile can be 256 queries wide and a batch of 1024 queries costs ONE pass
//     over the corpus in HBM instead of 32;
//   * the score matrix is still never written.  A strided SAMPLE of row tiles is scored first
//     (same kernel, SAMPLE = true: the epilogue reduces every 32-row group to its per-query
//     maximum); the r-th largest group maximum of a query is a lower bound of its r-th best score
//     over the whole corpus, so the main pass only has to APPEND the few thousand rows per query
//     that beat it to a per-query candidate buffer (staged per warp, flushed in bulk);
//   * operands are either the fp32 corpus read as tf32 (kind::tf32) or a bf16 shadow copy of it
//     (kind::f16, half the bytes, twice the rate, 3.2x wider error margin).
// Tensor-core scores only NOMINATE: dense_tc_rescore_kernel recomputes every candidate within a
// rigorous error margin of the k-th best in exact fp32 and flags the query for the exact scan
// when the lists cannot prove exactness (anr_dense_tc.cu).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "anr_internal.h"
#include "anr_tc.cuh"
#include "anr_topk.cuh"

namespace anr {

constexpr int kGmRows = 128;          // corpus rows per tile = UMMA M
constexpr int kGmThreads = 320;       // TMA producer warp, MMA warp, up to 8 epilogue warps
constexpr int kGmEpiWarps = 8;        // 4 or 8 launched: one or two per TMEM lane quadrant (two:
                                      // each takes every other 32-column chunk)
constexpr int kGmMaxStages = 12;
constexpr int kGmABytes = kGmRows * 128;   // one corpus slab: 128 rows x 128 bytes of K
constexpr int kGmThrThreads = 512;

struct GemmLayout {
  int thr_off, stage_off, bar_off, total_bytes;
  int n_stages, n_slabs, stage_bytes;
  int a_tiled;   // the corpus operand is the TILED bf16 shadow copy (see f32_to_bf16_tiled_kernel)
};

template <int NQ, bool BF16>
struct GemmCfg {
  static constexpr int kSlabElems = BF16 ? 64 : 32;            // elements per 128-byte K slab
  static constexpr int kBBytes = NQ * 128;
  static constexpr int kStageBytes = kGmABytes + kBBytes;
  static constexpr uint32_t kFmt = BF16 ? 1u : 2u;             // F16F32 format: 1 = bf16, 2 = tf32
  // D = F32 (bits 4-5 = 1), A/B format (bits 7-9, 10-12), both K-major, N >> 3 at 17, M >> 4 at 24
  static constexpr uint32_t kIdesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) |
                                     (static_cast<uint32_t>(NQ >> 3) << 17) |
                                     (static_cast<uint32_t>(kGmRows >> 4) << 24);
  static constexpr int kTmemCols = 2 * NQ;                     // 128, 256 or 512: powers of two
};

// element `lane` of v[] maximised over the 32 lanes (31 shuffles)
__device__ __forceinline__ float warp_transpose_max32(float (&v)[32], int lane) {
  int o = 16;
#pragma unroll
  for (int c = 32; c > 1; c >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < c / 2; ++j) {
      const float send = upper ? v[j] : v[j + c / 2];
      const float keep = upper ? v[j + c / 2] : v[j];
      v[j] = fmaxf(keep, __shfl_xor_sync(kFullMask, send, o));
    }
    o >>= 1;
  }
  return v[0];
}

// Survivors are staged per epilogue warp in shared memory (ballot + prefix count, no atomics) and
// appended to the per-query global buffers in bulk, many independent atomicAdds in flight at
// once: a returning global atomicAdd costs ~1 us, and issued one lane at a time from divergent
// branches they made the epilogue -- not the tensor pipe -- the bottleneck (profiles/).
constexpr int kGmStage = 192;        // staged survivors per epilogue warp
constexpr int kGmFlushAt = 96;       // flush at a chunk boundary once this many are staged
struct GemmStage {
  uint64_t key[kGmStage];
  int32_t q[kGmStage];
};

// A chunk's survivors (bit j of hm = column q0 + j beats its threshold) -> the warp's staging
// buffer; out of line: the hot loop only builds the masks (the caller spills v[] to its stack
// frame for the call, which is cheaper than a 31-select multiplexer per survivor).  Returns the
// new staged count.
static __device__ __noinline__ int gemm_stage_hits(GemmStage* st, int wn, uint32_t hm,
                                                   const uint32_t* v, int q0, uint32_t row,
                                                   uint64_t* __restrict__ cand,
                                                   int32_t* __restrict__ cnt, int cap, int lane);

// whole warp, converged; n = staged entries
static __device__ __noinline__ void gemm_flush(GemmStage* st, int n, uint64_t* __restrict__ cand,
                                               int32_t* __restrict__ cnt, int cap, int lane) {
  __syncwarp();
#pragma unroll 4
  for (int i = lane; i < n; i += 32) {
    const int q = st->q[i];
    const uint64_t key = st->key[i];
    const int slot = atomicAdd(cnt + q, 1);
    if (slot < cap) cand[static_cast<int64_t>(q) * cap + slot] = key;
  }
  __syncwarp();
}

static __device__ __noinline__ int gemm_stage_hits(GemmStage* st, int wn, uint32_t hm,
                                                   const uint32_t* v, int q0, uint32_t row,
                                                   uint64_t* __restrict__ cand,
                                                   int32_t* __restrict__ cnt, int cap, int lane) {
  const int mine = __popc(hm);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(kFullMask, incl, 31);
  if (total > kGmStage) {  // a burst (e.g. a block of duplicate rows): append directly
    while (hm) {
      const int j = __ffs(hm) - 1;
      hm &= hm - 1;
      const int q = q0 + j;
      const int slot = atomicAdd(cnt + q, 1);
      if (slot < cap)
        cand[static_cast<int64_t>(q) * cap + slot] = make_key(__uint_as_float(v[j]), row);
    }
    return wn;
  }
  if (wn + total > kGmStage) {
    gemm_flush(st, wn, cand, cnt, cap, lane);
    wn = 0;
  }
  int slot = wn + incl - mine;
  while (hm) {
    const int j = __ffs(hm) - 1;
    hm &= hm - 1;
    st->key[slot] = make_key(__uint_as_float(v[j]), row);
    st->q[slot] = q0 + j;
    ++slot;
  }
  return wn + total;
}

// Row tile t of this launch covers corpus rows [t * tile_stride * 128, +128).
//   SAMPLE = false: rows whose score beats thr[query] are appended to cand[query][cap] / cnt[query]
//   SAMPLE = true : gmax[query * gmax_stride + t * 4 + quadrant] = max score of that 32-row group
//
// PAIR = true: clusters of two CTAs on neighbouring row tiles share every QUERY slab -- rank r
// loads half of it and the TMA multicasts that half into both CTAs' rings, so the L2 -> SM
// traffic per 128 x NQ tile drops from 16 + NQ/8 KB to 16 + NQ/16 KB per slab.  (ncu: at 256
// queries per tile the single-CTA kernel is bound by the ~12 TB/s the L2 can deliver, not by the
// tensor pipe.)  A stage may be refilled only when BOTH CTAs' MMAs have read it: every commit
// arrives on the "empty" barrier of both CTAs.
template <int NQ, bool BF16, bool SAMPLE, bool PAIR>
__global__ void __launch_bounds__(kGmThreads, 1)   // launched with 64 + 32 * {4, 8} threads
dense_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  int64_t n, int64_t n_row_tiles, int64_t tile_stride, int n_qblocks,
                  const uint32_t* __restrict__ mask, const float* __restrict__ thr,
                  uint64_t* __restrict__ cand, int32_t* __restrict__ cnt, int cap,
                  float* __restrict__ gmax, int64_t gmax_stride, GemmLayout L,
                  int32_t* __restrict__ gate) {
  using Cfg = GemmCfg<NQ, BF16>;
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_wait();
  if (SAMPLE) pdl_trigger();   // (the long main pass triggers at its end: nothing parks beside it)
  if (gate && threadIdx.x == 0) atomicAdd(gate, 1);   // this CTA holds its shared memory now
  unsigned char* ring = smem;                                   // n_stages x [A slab | B slab]
  float* thr_s = reinterpret_cast<float*>(smem + L.thr_off);    // [n_qblocks * NQ]
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem + L.stage_off);   // [4 epilogue warps]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty = full + kGmMaxStages;
  uint64_t* acc_full = empty + kGmMaxStages;    // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_epi = static_cast<int>(blockDim.x >> 5) - 2;   // 4 or 8
  // work units: row tiles (PAIR: pairs of neighbouring row tiles, one per CTA of the cluster)
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int64_t unit0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int64_t unit_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int64_t n_units = PAIR ? (n_row_tiles + 1) / 2 : n_row_tiles;
  const int64_t my_row_tiles = n_units > unit0 ? (n_units - unit0 + unit_step - 1) / unit_step : 0;
  auto tile_of = [&](int64_t it) -> int64_t {
    const int64_t u = unit0 + it * unit_step;
    return PAIR ? 2 * u + rank : u;   // may be == n_row_tiles for the odd tile out: all zero rows
  };

  if (!SAMPLE)
    for (int i = threadIdx.x; i < n_qblocks * NQ; i += blockDim.x) thr_s[i] = thr[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < L.n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], PAIR ? 2 : 1);   // PAIR: one commit from each CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], n_epi);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer ----
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_row_tiles; ++it) {
        const int64_t tile_a = tile_of(it) * tile_stride;
        const int row0 = static_cast<int>(tile_a * kGmRows);
        // tiled shadow: slab kb of row tile t is the 16 KB block (t * n_slabs + kb)
        const int blk0 = static_cast<int>(tile_a * L.n_slabs * kGmRows);
        for (int qb = 0; qb < n_qblocks; ++qb) {
          for (int kb = 0; kb < L.n_slabs; ++kb) {
            mbar_wait(&empty[s], ph ^ 1u);
            unsigned char* st = ring + static_cast<size_t>(s) * Cfg::kStageBytes;
            mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
            if (L.a_tiled) tma_load_2d(st, &map_a, 0, blk0 + kb * kGmRows, &full[s]);
            else tma_load_2d(st, &map_a, kb * Cfg::kSlabElems, row0, &full[s]);
            if (PAIR)   // my half of the query slab, into both CTAs
              tma_load_2d_mcast(st + kGmABytes + rank * (Cfg::kBBytes / 2), &map_b,
                                kb * Cfg::kSlabElems, qb * NQ + rank * (NQ / 2), &full[s], 0x3);
            else
              tma_load_2d(st + kGmABytes, &map_b, kb * Cfg::kSlabElems, qb * NQ, &full[s]);
            if (++s == L.n_stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int64_t my_tiles = my_row_tiles * n_qblocks;
      for (int64_t t = 0; t < my_tiles; ++t) {
        const int a = static_cast<int>(t & 1);
        const uint32_t aph = static_cast<uint32_t>(t >> 1) & 1u;
        mbar_wait(&acc_empty[a], aph ^ 1u);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(a * NQ);
        for (int kb = 0; kb < L.n_slabs; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(ring + static_cast<size_t>(s) * Cfg::kStageBytes);
          const uint64_t da = tc_smem_desc(st);
          const uint64_t db = tc_smem_desc(st + kGmABytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 32 bytes of K per instruction (8 tf32 / 16 bf16)
            if (BF16)
              tc_mma_bf16(tmem_d, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2),
                          Cfg::kIdesc, (kb | kk) != 0 ? 1u : 0u);
            else
              tc_mma_tf32(tmem_d, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2),
                          Cfg::kIdesc, (kb | kk) != 0 ? 1u : 0u);
          }
          // the stage may be refilled once these MMAs (PAIR: and the peer's) have read it
          if (PAIR) tc_commit_mcast(&empty[s], 0x3);
          else tc_commit(&empty[s]);
          if (++s == L.n_stages) { s = 0; ph ^= 1u; }
        }
        tc_commit(&acc_full[a]);
      }
    }
  } else {
    // ---- epilogue: TMEM lane = corpus row, column = query ----
    const int quad = warp & 3;         // the TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;  // the two warps of a quadrant take alternate 32-column chunks
    GemmStage* st = stages + (warp - 2);
    int wn = 0;   // survivors staged by this warp (warp-uniform)
    int64_t t = 0;
    for (int64_t it = 0; it < my_row_tiles; ++it) {
      const int64_t tile = tile_of(it);
      const int64_t row = tile * tile_stride * kGmRows + quad * 32 + lane;
      bool ok = row < n && tile < n_row_tiles;
      if (ok && mask) ok = (__ldg(mask + (row >> 5)) >> (row & 31)) & 1u;
      for (int qb = 0; qb < n_qblocks; ++qb, ++t) {
        const int a = static_cast<int>(t & 1);
        const uint32_t aph = static_cast<uint32_t>(t >> 1) & 1u;
        mbar_wait(&acc_full[a], aph);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(a * NQ);
#pragma unroll 1
        for (int c = half; c < NQ / 32; c += (n_epi >> 2)) {
          uint32_t v[32];
          tc_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
          const int q0 = qb * NQ + c * 32;
          if (SAMPLE) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = ok ? __uint_as_float(v[j]) : -INFINITY;
            const float m = warp_transpose_max32(f, lane);
            if (tile < n_row_tiles)
              gmax[static_cast<int64_t>(q0 + lane) * gmax_stride + tile * 4 + quad] = m;
          } else {
            const float4* th = reinterpret_cast<const float4*>(thr_s + q0);
            uint32_t hm = 0u;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 tq = th[j4];
              hm |= (__uint_as_float(v[j4 * 4 + 0]) > tq.x ? 1u : 0u) << (j4 * 4 + 0);
              hm |= (__uint_as_float(v[j4 * 4 + 1]) > tq.y ? 1u : 0u) << (j4 * 4 + 1);
              hm |= (__uint_as_float(v[j4 * 4 + 2]) > tq.z ? 1u : 0u) << (j4 * 4 + 2);
              hm |= (__uint_as_float(v[j4 * 4 + 3]) > tq.w ? 1u : 0u) << (j4 * 4 + 3);
            }
            if (!ok) hm = 0u;
            if (__any_sync(kFullMask, hm != 0u))
              wn = gemm_stage_hits(st, wn, hm, v, q0, static_cast<uint32_t>(row), cand, cnt, cap, lane);
            if (wn >= kGmFlushAt) {
              gemm_flush(st, wn, cand, cnt, cap, lane);
              wn = 0;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[a]);
      }
    }
    if (!SAMPLE) gemm_flush(st, wn, cand, cnt, cap, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (!SAMPLE) pdl_trigger();
  if (PAIR) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

// ---- cta_group::2: one 256-row x NQ-query MMA over a CTA pair ------------------------------------
// The CTAs of a cluster of two own neighbouring 128-row tiles.  Each keeps ITS rows of the corpus
// slab and ITS HALF of the query slab (NQ / 2 rows) in shared memory; the leader CTA (rank 0)
// issues tcgen05.mma.cta_group::2 (M = 256): the tensor cores of both SMs read A from their own
// CTA and the two halves of B from both, and write each CTA's 128 x NQ accumulator into its own
// TMEM.  Per CTA and K slab that is 16 + NQ/16 KB of TMA writes and operand reads instead of
// 16 + NQ/8 KB -- the shared-memory traffic that bounds the single-CTA kernel at NQ = 256.
//   full[s]      (leader only) armed by the leader's producer with the bytes of BOTH CTAs; the
//                peer's TMA loads signal it through its shared::cluster address
//   empty[s]     (both) one multicast commit from the leader's MMA thread
//   acc_full[a]  (both) one multicast commit;  acc_empty[a] (leader) all epilogue warps of both CTAs
template <int NQ, bool BF16, bool SAMPLE>
__global__ void __launch_bounds__(kGmThreads, 1)
dense_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   int64_t n, int64_t n_row_tiles, int64_t tile_stride, int n_qblocks,
                   const uint32_t* __restrict__ mask, const float* __restrict__ thr,
                   uint64_t* __restrict__ cand, int32_t* __restrict__ cnt, int cap,
                   float* __restrict__ gmax, int64_t gmax_stride, GemmLayout L) {
  using Cfg = GemmCfg<NQ, BF16>;
  constexpr int kBHalf = Cfg::kBBytes / 2;
  constexpr int kStage2 = kGmABytes + kBHalf;
  // M = 256 (both CTAs), N = NQ
  constexpr uint32_t kIdesc2 = (1u << 4) | (Cfg::kFmt << 7) | (Cfg::kFmt << 10) |
                               (static_cast<uint32_t>(NQ >> 3) << 17) | (16u << 24);
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_wait();
  if (SAMPLE) pdl_trigger();
  unsigned char* ring = smem;
  float* thr_s = reinterpret_cast<float*>(smem + L.thr_off);
  GemmStage* stages = reinterpret_cast<GemmStage*>(smem + L.stage_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty = full + kGmMaxStages;
  uint64_t* acc_full = empty + kGmMaxStages;    // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_epi = static_cast<int>(blockDim.x >> 5) - 2;
  const int rank = static_cast<int>(cluster_ctarank());
  const int64_t unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;
  const int64_t n_units = (n_row_tiles + 1) / 2;
  const int64_t my_units = n_units > unit0 ? (n_units - unit0 + unit_step - 1) / unit_step : 0;
  auto tile_of = [&](int64_t it) -> int64_t { return 2 * (unit0 + it * unit_step) + rank; };

  if (!SAMPLE)
    for (int i = threadIdx.x; i < n_qblocks * NQ; i += blockDim.x) thr_s[i] = thr[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < L.n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 2 * n_epi);
    }
    mbar_fence_init();
  }
  if (warp == 1) {   // one warp of EACH CTA: the pair allocates the same columns in both TMEMs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer (both CTAs): my rows of A, my half of B; bytes counted on the leader ----
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_units; ++it) {
        const int64_t tile_a = tile_of(it) * tile_stride;
        const int row0 = static_cast<int>(tile_a * kGmRows);
        const int blk0 = static_cast<int>(tile_a * L.n_slabs * kGmRows);   // (tiled shadow)
        for (int qb = 0; qb < n_qblocks; ++qb) {
          for (int kb = 0; kb < L.n_slabs; ++kb) {
            mbar_wait(&empty[s], ph ^ 1u);
            unsigned char* st = ring + static_cast<size_t>(s) * kStage2;
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * kStage2);
            const uint32_t lead_full = mapa_shared(smem_u32(&full[s]), 0);
            if (L.a_tiled) tma_load_2d_cg2(st, &map_a, 0, blk0 + kb * kGmRows, lead_full);
            else tma_load_2d_cg2(st, &map_a, kb * Cfg::kSlabElems, row0, lead_full);
            tma_load_2d_cg2(st + kGmABytes, &map_b, kb * Cfg::kSlabElems, qb * NQ + rank * (NQ / 2),
                            lead_full);
            if (++s == L.n_stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: the leader CTA only ----
    if (lane == 0 && rank == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int64_t my_tiles = my_units * n_qblocks;
      for (int64_t t = 0; t < my_tiles; ++t) {
        const int a = static_cast<int>(t & 1);
        const uint32_t aph = static_cast<uint32_t>(t >> 1) & 1u;
        mbar_wait(&acc_empty[a], aph ^ 1u);  // the epilogues of BOTH CTAs have drained it
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(a * NQ);
        for (int kb = 0; kb < L.n_slabs; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(ring + static_cast<size_t>(s) * kStage2);
          const uint64_t da = tc_smem_desc(st);
          const uint64_t db = tc_smem_desc(st + kGmABytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (BF16)
              tc2_mma_bf16(tmem_d, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2),
                           kIdesc2, (kb | kk) != 0 ? 1u : 0u);
            else
              tc2_mma_tf32(tmem_d, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2),
                           kIdesc2, (kb | kk) != 0 ? 1u : 0u);
          }
          tc2_commit_mcast(&empty[s], 0x3);
          if (++s == L.n_stages) { s = 0; ph ^= 1u; }
        }
        tc2_commit_mcast(&acc_full[a], 0x3);
      }
    }
  } else {
    // ---- epilogue (both CTAs): as in dense_gemm_kernel ----
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    GemmStage* st = stages + (warp - 2);
    int wn = 0;
    int64_t t = 0;
    for (int64_t it = 0; it < my_units; ++it) {
      const int64_t tile = tile_of(it);
      const int64_t row = tile * tile_stride * kGmRows + quad * 32 + lane;
      bool ok = row < n && tile < n_row_tiles;
      if (ok && mask) ok = (__ldg(mask + (row >> 5)) >> (row & 31)) & 1u;
      for (int qb = 0; qb < n_qblocks; ++qb, ++t) {
        const int a = static_cast<int>(t & 1);
        const uint32_t aph = static_cast<uint32_t>(t >> 1) & 1u;
        mbar_wait(&acc_full[a], aph);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(a * NQ);
#pragma unroll 1
        for (int c = half; c < NQ / 32; c += (n_epi >> 2)) {
          uint32_t v[32];
          tc_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
          const int q0 = qb * NQ + c * 32;
          if (SAMPLE) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = ok ? __uint_as_float(v[j]) : -INFINITY;
            const float m = warp_transpose_max32(f, lane);
            if (tile < n_row_tiles)
              gmax[static_cast<int64_t>(q0 + lane) * gmax_stride + tile * 4 + quad] = m;
          } else {
            const float4* th = reinterpret_cast<const float4*>(thr_s + q0);
            uint32_t hm = 0u;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 tq = th[j4];
              hm |= (__uint_as_float(v[j4 * 4 + 0]) > tq.x ? 1u : 0u) << (j4 * 4 + 0);
              hm |= (__uint_as_float(v[j4 * 4 + 1]) > tq.y ? 1u : 0u) << (j4 * 4 + 1);
              hm |= (__uint_as_float(v[j4 * 4 + 2]) > tq.z ? 1u : 0u) << (j4 * 4 + 2);
              hm |= (__uint_as_float(v[j4 * 4 + 3]) > tq.w ? 1u : 0u) << (j4 * 4 + 3);
            }
            if (!ok) hm = 0u;
            if (__any_sync(kFullMask, hm != 0u))
              wn = gemm_stage_hits(st, wn, hm, v, q0, static_cast<uint32_t>(row), cand, cnt, cap, lane);
            if (wn >= kGmFlushAt) {
              gemm_flush(st, wn, cand, cnt, cap, lane);
              wn = 0;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[a]), 0));
      }
    }
    if (!SAMPLE) gemm_flush(st, wn, cand, cnt, cap, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (!SAMPLE) pdl_trigger();
  cluster_sync_all();   // both CTAs are done with each other's shared memory, barriers and TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

// Starting thresholds from the sample's group maxima.  The rank-th largest of the 512 per-thread
// bests is reached by `rank` distinct groups, hence by `rank` distinct rows: a lower bound of the
// rank-th best score of the corpus.  Queries >= n_real (zero padding) never nominate anything.
__global__ void __launch_bounds__(kGmThrThreads)
dense_gemm_thr_kernel(const float* __restrict__ gmax, int64_t gmax_stride, int m, int rank,
                      int n_real, float* __restrict__ thr, uint64_t* __restrict__ thr_key,
                      int32_t* __restrict__ cnt, int32_t* __restrict__ gate,
                      int32_t* __restrict__ ticket) {
  __shared__ uint64_t best[kGmThrThreads];
  pdl_wait();
  pdl_trigger();
  const int q = blockIdx.x;
  if (threadIdx.x == 0) {   // the main pass starts from empty candidate buffers and a closed gate
    cnt[q] = 0;
    if (q == 0 && gate) *gate = 0;
    if (q == 0 && ticket) *ticket = 0;   // (the rescoring kernel's last-CTA ticket)
  }
  if (q >= n_real) {
    if (threadIdx.x == 0) {
      thr[q] = INFINITY;
      thr_key[q] = 0ull;
    }
    return;
  }
  const float* g = gmax + static_cast<int64_t>(q) * gmax_stride;
  float b = -INFINITY;
  for (int i = threadIdx.x; i < m; i += kGmThrThreads) b = fmaxf(b, g[i]);
  __shared__ uint64_t s_kth;
  const uint64_t kth = block_kth_of_thread_bests<kGmThrThreads / 32>(
      b == -INFINITY ? 0ull : make_key(b, 0u), rank, best, &s_kth);
  if (threadIdx.x == 0) {
    thr[q] = kth ? key_score(kth) : -INFINITY;
    thr_key[q] = kth ? (kth | 0xffffffffull) : 0ull;
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                   int64_t n4) {
  pdl_wait();
  pdl_trigger();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 p;
    p.x = *reinterpret_cast<uint32_t*>(&lo);
    p.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = p;
  }
}

cudaError_t launch_f32_to_bf16(const float* in, void* out, int64_t count, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  const int64_t n4 = count / 4;  // callers pass multiples of 4 (ld % 4 == 0)
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  return launch_chain(f32_to_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream,
                      in, static_cast<__nv_bfloat16*>(out), n4);
}

// The bf16 shadow copy of the corpus in TILED order: the 128-row x 64-column block (t, kb) --
// what one ring stage of the GEMM pass holds -- is one contiguous 16 KB run at block index
// t * n_slabs + kb, rows of 128 bytes.  A stage is then ONE sequential 16 KB read instead of 128
// pieces of 128 bytes 2 KB apart (the row-major copy: 6.25-6.7 TB/s for the 64-query pass against
// 7.4 TB/s for the fp32 scan's sequential 16 KB bulk copies).  Rows >= n of the last tile are zero.
__global__ void f32_to_bf16_tiled_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                         int64_t n, int ld4, int n_slabs, int64_t total4) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / ld4;
    const int col = static_cast<int>(i - row * ld4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) v = __ldg(reinterpret_cast<const float4*>(in) + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 p;
    p.x = *reinterpret_cast<uint32_t*>(&lo);
    p.y = *reinterpret_cast<uint32_t*>(&hi);
    const int64_t t = row / kGmRows;
    const int r = static_cast<int>(row - t * kGmRows);
    const int64_t blk = t * n_slabs + (col >> 6);
    const int64_t dst = (blk * kGmRows + r) * 64 + (col & 63);   // in elements
    reinterpret_cast<uint2*>(out)[dst >> 2] = p;
  }
}

static int g_pdl_mask = getenv("ANR_PDL") ? atoi(getenv("ANR_PDL")) : 7;
bool pdl_enabled(int which) { return (g_pdl_mask & which) == which; }
void pdl_set_mask(int mask) { g_pdl_mask = mask; }

static bool shadow_tiled_env() {
  static const bool on = !(getenv("ANR_SHADOW_TILED") && atoi(getenv("ANR_SHADOW_TILED")) == 0);
  return on;
}
// Layout of the shadow copy of an [n, ld] corpus (ld % 64 == 0): tiled unless ANR_SHADOW_TILED=0 or
// the block rows would overflow the TMA's 32-bit coordinates.
bool dense_shadow_tiled(int64_t n, int ld) {
  const int64_t n_tiles = (n + kGmRows - 1) / kGmRows;
  return shadow_tiled_env() && ld % 64 == 0 && (n_tiles + 1) * (ld / 64) * kGmRows < (1ll << 31);
}
size_t dense_shadow_bytes(int64_t n, int ld) {
  const int64_t rows = dense_shadow_tiled(n, ld) ? (n + kGmRows - 1) / kGmRows * kGmRows : n;
  return static_cast<size_t>(rows) * ld * 2;
}
cudaError_t launch_dense_shadow_fill(const float* emb, void* shadow, int64_t n, int ld,
                                     cudaStream_t stream) {
  if (!dense_shadow_tiled(n, ld)) return launch_f32_to_bf16(emb, shadow, n * ld, stream);
  const int64_t rows = (n + kGmRows - 1) / kGmRows * kGmRows;
  const int64_t total4 = rows * (ld / 4);
  int64_t blocks = (total4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_bf16_tiled_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      emb, static_cast<__nv_bfloat16*>(shadow), n, ld / 4, ld / 64, total4);
  return cudaGetLastError();
}

// One thread polls the gate (DenseGate, anr_internal.h); gives up after ~10 ms so that a dense
// pass that never starts (an error upstream) cannot hang the BM25 stream.
__global__ void gate_wait_kernel(const int32_t* __restrict__ counter, int expected) {
  const long long t0 = clock64();
  while (*reinterpret_cast<const volatile int32_t*>(counter) < expected) {
    __nanosleep(200);
    if (clock64() - t0 > 20000000ll) break;
  }
}
cudaError_t launch_gate_wait(const int32_t* counter, int expected, cudaStream_t stream) {
  if (!counter || expected < 1) return cudaSuccess;
  gate_wait_kernel<<<1, 1, 0, stream>>>(counter, expected);
  return cudaGetLastError();
}

// ---- host side ---------------------------------------------------------------------------
constexpr int kGmTargetCand = 4096;   // rows per query the sample threshold should let through
constexpr int kGmCap = 4 * kGmTargetCand;

static int gemm_thr_rank(int k) { return k <= 10 ? 16 : k + 32; }
// Rank actually used for a corpus of n rows sampled by sample_tiles tiles.  When the sample is a
// large part of a small corpus (one wave of tiles is the minimum), rank 16 of the sample sits only
// ~100 rows deep in the corpus: inside the bf16 error margin of the k-th best, so most queries
// were flagged for the exact rescan (125k-row shards: 0.73 ms per 64 queries instead of 0.2).
// The threshold is therefore kept at least kGmMinCand corpus rows deep.
constexpr int kGmMinCand = 1024;
static int gemm_thr_rank_for(int64_t n, int64_t sample_tiles, int k) {
  const int64_t sample_rows = sample_tiles * kGmRows;
  const int64_t deep = (kGmMinCand * sample_rows + n - 1) / n;
  int64_t rank = std::max<int64_t>(gemm_thr_rank(k), deep);
  rank = std::min<int64_t>(rank, 448);                       // <= per-thread bests of the kernel
  rank = std::min<int64_t>(rank, sample_tiles * 4);          // <= number of group maxima
  return static_cast<int>(std::max<int64_t>(rank, 1));
}

// rows of the strided sample: rank * n / sample ~= kGmTargetCand, whole waves of tiles
static int64_t gemm_sample_tiles(const DeviceProps& dp, int64_t n, int k) {
  const int64_t want_rows = (static_cast<int64_t>(gemm_thr_rank(k)) * n + kGmTargetCand - 1) / kGmTargetCand;
  int64_t tiles = (want_rows + kGmRows - 1) / kGmRows;
  tiles = (tiles + dp.sm_count - 1) / dp.sm_count * dp.sm_count;
  return tiles;
}

int dense_gemm_block(int nq) { return nq <= 64 ? 64 : (nq <= 128 ? 128 : 256); }
int dense_gemm_max_queries() { return 1024; }
int dense_gemm_padded_queries(int nq) {
  const int b = dense_gemm_block(nq);
  return (nq + b - 1) / b * b;
}

// b_rows: query rows each CTA holds per slab (nq_block, or nq_block / 2 under cta_group::2)
// epilogue warps a launch group of nq_block-query tiles runs with (one or two per TMEM quadrant)
static int gemm_epi_warps(int nq_block) {
  static const int epi_env = getenv("ANR_GEMM_EPI_WARPS") ? atoi(getenv("ANR_GEMM_EPI_WARPS")) : 0;
  return epi_env == 4 || epi_env == 8 ? epi_env : (nq_block >= 128 ? 8 : 4);
}

// Shared memory (bytes) the GEMM launches issued by this thread leave free on every SM: the short
// kernels of the BM25 path that runs beside the dense pass of a hybrid step (anr_api.cu) hold a few
// KB of static shared memory each -- its rerun kernels 34 + 18 KB -- and a dense CTA that takes the
// whole SM makes them wait until it exits (measured: the 2 KB ranking kernel of the BM25 path sat
// 218 us behind the 9-stage ring of the 64-query pass, profiles/r2_call24_*).
static thread_local int t_leave_smem = 0;
void dense_gemm_set_leave_smem(int bytes) { t_leave_smem = bytes > 0 ? bytes : 0; }

static bool make_gemm_layout(const DeviceProps& dp, int ld, bool bf16, int b_rows, int nq_pad,
                             GemmLayout* L, int ring_cap = 0, int n_epi = kGmEpiWarps) {
  const int slab = bf16 ? 64 : 32;
  if (ld % slab != 0) return false;
  L->n_slabs = ld / slab;
  L->stage_bytes = kGmABytes + b_rows * 128;
  const int thr_bytes = nq_pad * 4;
  const int bar_bytes = (2 * kGmMaxStages + 4) * 8 + 16;
  const int stage_bytes = n_epi * static_cast<int>(sizeof(GemmStage));   // one staging buffer per warp
  const int avail =
      dp.max_smem_optin - 1024 /* alignment slack */ - thr_bytes - stage_bytes - bar_bytes - 256;
  int n_stages = avail / L->stage_bytes;
  if (n_stages < 3) return false;
  if (n_stages > kGmMaxStages) n_stages = kGmMaxStages;
  if (t_leave_smem > 0) {   // (never below 4 stages: the ring has to cover the HBM latency)
    const int fit = (avail - t_leave_smem) / L->stage_bytes;
    n_stages = std::min(n_stages, std::max(fit, std::min(n_stages, 4)));
  }
  // a shorter ring leaves shared memory to the BM25 CTAs of a hybrid step (gemm_ring_cap)
  if (ring_cap >= 3 && n_stages > ring_cap) n_stages = ring_cap;
  L->n_stages = n_stages;
  L->thr_off = n_stages * L->stage_bytes;
  L->stage_off = (L->thr_off + thr_bytes + 15) / 16 * 16;
  L->bar_off = L->stage_off + stage_bytes;
  L->total_bytes = L->bar_off + bar_bytes;
  return true;
}

bool dense_gemm_supported(const DeviceProps& dp, int64_t n, int ld, int k, bool bf16) {
  if (getenv("ANR_DISABLE_GEMM") || getenv("ANR_DISABLE_TC")) return false;
  if (k < 1 || k > 128 || n >= (1ll << 31)) return false;
  GemmLayout L;
  if (!make_gemm_layout(dp, ld, bf16, 256, dense_gemm_max_queries(), &L)) return false;
  // the sample must be a small part of the corpus, or the pre-pass does not pay off
  const int64_t n_tiles = (n + kGmRows - 1) / kGmRows;
  return n_tiles >= 4 * gemm_sample_tiles(dp, n, k);
}

// scratch (bytes) one launch of up to dense_gemm_max_queries() queries needs
size_t dense_gemm_scratch_bytes(const DeviceProps& dp, int64_t n, int ld, int nq, int k) {
  const size_t nq_pad = static_cast<size_t>(dense_gemm_padded_queries(std::min(nq, dense_gemm_max_queries())));
  const size_t groups = static_cast<size_t>(gemm_sample_tiles(dp, n, k)) * 4;
  return nq_pad * kGmCap * 8 + nq_pad * groups * 4 + nq_pad * (4 + 4 + 8) +
         nq_pad * static_cast<size_t>(ld) * 2 + 4096 + 512;
}

// host-side only (GemmLayout is a kernel parameter and stays as it is): the launch group being
// issued by this thread runs with a capped ring
static thread_local bool t_ring_capped = false;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn gemm_encode_fn() {
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
// 2-D [rows, ld] row-major matrix of fp32 or bf16, box [box_rows x 128 bytes], 128-byte swizzle
static bool gemm_encode_map(CUtensorMap* map, const void* base, int64_t rows, int ld, int box_rows,
                            bool bf16) {
  EncodeTiledFn fn = gemm_encode_fn();
  if (!fn) return false;
  const int esz = bf16 ? 2 : 4;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
            const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the tiled shadow copy as a 2-D tensor of 128-byte rows: box = one 16 KB block
static bool gemm_encode_map_tiled(CUtensorMap* map, const void* base, int64_t n, int ld) {
  EncodeTiledFn fn = gemm_encode_fn();
  if (!fn) return false;
  const int64_t n_tiles = (n + kGmRows - 1) / kGmRows;
  const cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(n_tiles * (ld / 64) * kGmRows)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kGmRows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NQ, bool BF16, bool SAMPLE, bool PAIR>
static cudaError_t gemm_launch_one(int grid, int threads, int smem, cudaStream_t stream,
                                   const CUtensorMap& map_a, const CUtensorMap& map_b, int64_t n,
                                   int64_t n_row_tiles, int64_t tile_stride, int n_qblocks,
                                   const uint32_t* mask, const float* thr, uint64_t* cand,
                                   int32_t* cnt, int cap, float* gmax, int64_t gstride,
                                   const GemmLayout& L, int32_t* gate = nullptr, bool pdl = true) {
  auto kern = dense_gemm_kernel<NQ, BF16, SAMPLE, PAIR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  // a capped ring must keep the max-shared L1 split the BM25 kernels ask for, or CTAs of the two
  // kernels cannot share an SM (without this preference a 4-stage ring gained nothing: 0.651 ms)
  if (t_ring_capped)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl && pdl_enabled(SAMPLE ? kPdlDense : (kPdlDense | kPdlDenseMain)) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, map_a, map_b, n, n_row_tiles, tile_stride, n_qblocks, mask,
                            thr, cand, cnt, cap, gmax, gstride, L, gate);
}

template <int NQ, bool BF16, bool SAMPLE>
static cudaError_t gemm2_launch_one(int grid, int threads, int smem, cudaStream_t stream,
                                    const CUtensorMap& map_a, const CUtensorMap& map_b_half, int64_t n,
                                    int64_t n_row_tiles, int64_t tile_stride, int n_qblocks,
                                    const uint32_t* mask, const float* thr, uint64_t* cand,
                                    int32_t* cnt, int cap, float* gmax, int64_t gstride,
                                    const GemmLayout& L, bool pdl = true) {
  auto kern = dense_gemm2_kernel<NQ, BF16, SAMPLE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  if (t_ring_capped)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl && pdl_enabled(SAMPLE ? kPdlDense : (kPdlDense | kPdlDenseMain)) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, map_a, map_b_half, n, n_row_tiles, tile_stride, n_qblocks,
                            mask, thr, cand, cnt, cap, gmax, gstride, L);
}

// 0 = one CTA per tile, 1 = CTA pairs multicast the query slabs, 2 = cta_group::2 MMAs over a pair
static int gemm_mode(const DeviceProps& dp, int nq_block, int64_t n_tiles) {
  static const int env = getenv("ANR_GEMM_MODE") ? atoi(getenv("ANR_GEMM_MODE")) : -1;
  if (dp.sm_count % 2 != 0 || n_tiles < dp.sm_count) return 0;
  if (env >= 0 && env <= 2) return env;
  return nq_block >= 128 ? 2 : 0;
}

// sample pass + thresholds + main pass.  map_b_half: the query matrix with boxes of NQ / 2 rows
// (what each CTA of a pair loads and multicasts).
template <int NQ, bool BF16>
static cudaError_t gemm_launch_pair(const DeviceProps& dp, const CUtensorMap& map_a,
                                    const CUtensorMap& map_b, const CUtensorMap& map_b_half,
                                    int64_t n, int n_qblocks, const uint32_t* mask,
                                    int64_t sample_tiles, int n_real, int k, float* gmax, float* thr,
                                    uint64_t* thr_key, uint64_t* cand, int32_t* cnt,
                                    const GemmLayout& L, cudaEvent_t ev_start, cudaEvent_t ev_stop,
                                    cudaStream_t stream, cudaEvent_t ev_pre_main, int32_t* gate_ctr,
                                    DenseGate* gate) {
  const int smem = L.total_bytes + 1024;  // room to align the dynamic base to 1024 bytes
  const int64_t n_tiles = (n + kGmRows - 1) / kGmRows;
  const int64_t stride = n_tiles / sample_tiles;
  const int64_t gstride = sample_tiles * 4;
  const int nq_pad = n_qblocks * NQ;
  // two epilogue warps per quadrant once a tile has enough 32-column chunks to share
  const int n_epi = gemm_epi_warps(NQ);
  const int threads = 64 + 32 * n_epi;
  const int mode = gemm_mode(dp, NQ, n_tiles);
  const bool pair = mode == 1;
  cudaError_t e;
  if (mode == 2) {
    e = gemm2_launch_one<NQ, BF16, true>(dp.sm_count, threads, smem, stream, map_a, map_b_half, n,
                                         sample_tiles, stride, n_qblocks, mask, nullptr, nullptr,
                                         nullptr, 0, gmax, gstride, L);
    if (e != cudaSuccess) return e;
    e = launch_chain(dense_gemm_thr_kernel, dim3(nq_pad), dim3(kGmThrThreads), 0, stream, gmax, gstride,
                     static_cast<int>(gstride), gemm_thr_rank_for(n, sample_tiles, k), n_real, thr,
                     thr_key, cnt, static_cast<int32_t*>(nullptr), gate_ctr + 8);
    if (e != cudaSuccess) return e;
    if (ev_pre_main) cudaEventRecord(ev_pre_main, stream);   // the main kernel is next in line
    if (ev_start) cudaEventRecord(ev_start, stream);
    e = gemm2_launch_one<NQ, BF16, false>(dp.sm_count, threads, smem, stream, map_a, map_b_half, n,
                                          n_tiles, 1, n_qblocks, mask, thr, cand, cnt, kGmCap, nullptr,
                                          0, L, ev_start == nullptr);
    if (ev_stop) cudaEventRecord(ev_stop, stream);
    return e != cudaSuccess ? e : cudaGetLastError();
  }
  if (pair)
    e = gemm_launch_one<NQ, BF16, true, true>(dp.sm_count, threads, smem, stream, map_a, map_b_half, n,
                                              sample_tiles, stride, n_qblocks, mask, nullptr, nullptr,
                                              nullptr, 0, gmax, gstride, L);
  else
    e = gemm_launch_one<NQ, BF16, true, false>(dp.sm_count, threads, smem, stream, map_a, map_b, n,
                                               sample_tiles, stride, n_qblocks, mask, nullptr, nullptr,
                                               nullptr, 0, gmax, gstride, L);
  if (e != cudaSuccess) return e;
  const int grid = static_cast<int>(std::min<int64_t>(dp.sm_count, n_tiles));
  const bool gated = gate != nullptr && !pair;
  e = launch_chain(dense_gemm_thr_kernel, dim3(nq_pad), dim3(kGmThrThreads), 0, stream, gmax, gstride,
                   static_cast<int>(gstride), gemm_thr_rank_for(n, sample_tiles, k), n_real, thr,
                   thr_key, cnt, gated ? gate_ctr : static_cast<int32_t*>(nullptr), gate_ctr + 8);
  if (e != cudaSuccess) return e;
  if (gated) {
    gate->counter = gate_ctr;
    gate->expected = grid;
  }
  if (ev_pre_main) cudaEventRecord(ev_pre_main, stream);   // the main kernel is next in line
  if (ev_start) cudaEventRecord(ev_start, stream);   // brackets the main GEMM kernel only
  if (pair)
    e = gemm_launch_one<NQ, BF16, false, true>(dp.sm_count, threads, smem, stream, map_a, map_b_half, n,
                                               n_tiles, 1, n_qblocks, mask, thr, cand, cnt, kGmCap,
                                               nullptr, 0, L, nullptr, ev_start == nullptr);
  else   // (events around the kernel, i.e. a profiled step: a plain, fully serialised launch)
    e = gemm_launch_one<NQ, BF16, false, false>(grid, threads, smem, stream, map_a, map_b, n, n_tiles,
                                                1, n_qblocks, mask, thr, cand, cnt, kGmCap, nullptr,
                                                0, L, gated ? gate_ctr : nullptr, ev_start == nullptr);
  if (ev_stop) cudaEventRecord(ev_stop, stream);
  return e != cudaSuccess ? e : cudaGetLastError();
}

// Ring depth cap of one launch group.  ANR_GEMM_MAX_STAGES=n caps every GEMM launch (profiling).
// ANR_GEMM_BESIDE_STAGES=n (default 4, 0 = off) caps the launches of a hybrid step whose BM25 scan
// runs beside the main kernel (64-query bf16 tiles, 16+ queries): with a 4-stage ring (96 KB)
// three BM25 CTAs fit next to the dense CTA on every SM, and the two main kernels -- one HBM-bound,
// one issue-bound -- overlap instead of following each other.  Measured on 1M x 1024 + 1M docs,
// batch 64: 0.629 -> 0.563 ms per step (round 2, profiles/r2_call1_*; the whole GPU suite passes
// with it), the dense kernel stretching from 0.32 to 0.39 ms; batch-1 loses 3 %, hence the
// 16-query floor.
static int gemm_ring_cap(bool beside_bm25, bool bf16, int nqb_size, int n_real) {
  static const int all_env = getenv("ANR_GEMM_MAX_STAGES") ? atoi(getenv("ANR_GEMM_MAX_STAGES")) : 0;
  static const int beside_env =
      getenv("ANR_GEMM_BESIDE_STAGES") ? atoi(getenv("ANR_GEMM_BESIDE_STAGES")) : 4;
  if (all_env >= 3) return all_env;
  if (beside_env >= 3 && beside_bm25 && bf16 && nqb_size == 64 && n_real >= 16) return beside_env;
  return 0;
}

// One launch group: n_real <= dense_gemm_max_queries() queries at q_dev ([padded, ld] fp32, zero
// rows as padding) against the corpus (emb fp32 [n, ld]; shadow = its bf16 copy or null ->
// tf32 on the fp32 words).  Exact top-k of the n_real queries through `out`, flags[0..n_real).
cudaError_t launch_dense_gemm(const DeviceProps& dp, const float* emb, const void* shadow, int64_t n,
                              int ld, const float* q_dev, int n_real, int k, const uint32_t* mask,
                              float emb_norm_max, unsigned char* scratch, const TopkOut& out,
                              int32_t* flags, cudaEvent_t ev_start, cudaEvent_t ev_stop,
                              cudaStream_t stream, cudaEvent_t ev_pre_main, DenseGate* gate,
                              int32_t* n_flagged, int32_t* flagged) {
  const bool bf16 = shadow != nullptr;
  const int nqb_size = dense_gemm_block(n_real);
  const int nq_pad = dense_gemm_padded_queries(n_real);
  const int n_qblocks = nq_pad / nqb_size;
  GemmLayout L;
  const int mode = gemm_mode(dp, nqb_size, (n + kGmRows - 1) / kGmRows);
  const int ring_cap = gemm_ring_cap(ev_pre_main != nullptr, bf16, nqb_size, n_real);
  t_ring_capped = ring_cap >= 3 || t_leave_smem > 0;   // (both keep the max-shared L1 split)
  if (!make_gemm_layout(dp, ld, bf16, mode == 2 ? nqb_size / 2 : nqb_size, nq_pad, &L, ring_cap,
                        gemm_epi_warps(nqb_size)))
    return cudaErrorInvalidConfiguration;
  L.a_tiled = bf16 && dense_shadow_tiled(n, ld) ? 1 : 0;
  const int64_t sample_tiles = gemm_sample_tiles(dp, n, k);

  // carve the scratch
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = scratch + off;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return p;
  };
  uint64_t* cand = reinterpret_cast<uint64_t*>(take(static_cast<size_t>(nq_pad) * kGmCap * 8));
  float* gmax = reinterpret_cast<float*>(take(static_cast<size_t>(nq_pad) * sample_tiles * 4 * 4));
  float* thr = reinterpret_cast<float*>(take(static_cast<size_t>(nq_pad) * 4));
  int32_t* cnt = reinterpret_cast<int32_t*>(take(static_cast<size_t>(nq_pad) * 4));
  uint64_t* thr_key = reinterpret_cast<uint64_t*>(take(static_cast<size_t>(nq_pad) * 8));
  int32_t* gate_ctr = reinterpret_cast<int32_t*>(take(256));
  const void* q_ops = q_dev;
  if (bf16) {
    void* q16 = take(static_cast<size_t>(nq_pad) * ld * 2);
    cudaError_t e = launch_f32_to_bf16(q_dev, q16, static_cast<int64_t>(nq_pad) * ld, stream);
    if (e != cudaSuccess) return e;
    q_ops = q16;
  }
  cudaError_t e = cudaSuccess;   // (cnt is zeroed by the threshold kernel, right before the main pass)

  CUtensorMap map_a, map_b, map_b_half;
  if (!(L.a_tiled ? gemm_encode_map_tiled(&map_a, shadow, n, ld)
                  : gemm_encode_map(&map_a, bf16 ? shadow : static_cast<const void*>(emb), n, ld,
                                    kGmRows, bf16)) ||
      !gemm_encode_map(&map_b, q_ops, nq_pad, ld, nqb_size, bf16) ||
      !gemm_encode_map(&map_b_half, q_ops, nq_pad, ld, nqb_size / 2, bf16))
    return cudaErrorInvalidValue;

#define ANR_GEMM_CASE(NQV, BFV)                                                                   \
  e = gemm_launch_pair<NQV, BFV>(dp, map_a, map_b, map_b_half, n, n_qblocks, mask, sample_tiles,  \
                                 n_real, k, gmax, thr, thr_key, cand, cnt, L, ev_start, ev_stop,   \
                                 stream, ev_pre_main, gate_ctr, gate)
  if (bf16) {
    if (nqb_size == 64) ANR_GEMM_CASE(64, true);
    else if (nqb_size == 128) ANR_GEMM_CASE(128, true);
    else ANR_GEMM_CASE(256, true);
  } else {
    if (nqb_size == 64) ANR_GEMM_CASE(64, false);
    else if (nqb_size == 128) ANR_GEMM_CASE(128, false);
    else ANR_GEMM_CASE(256, false);
  }
#undef ANR_GEMM_CASE
  if (e != cudaSuccess) return e;
  // error bound of the product sum relative to |q| * |e| (Cauchy-Schwarz): tf32 truncates each
  // operand by < 2^-10 relative; bf16 rounds to nearest, <= 2^-8 relative per operand; plus
  // fp32 accumulation noise
  const float eps_rel = bf16 ? 8.1e-3f : 2.5e-3f;
  // n_flagged != nullptr: this launch group is the whole batch, and the rescoring kernel's last CTA
  // lists the flagged queries itself (ticket zeroed by the threshold kernel)
  return launch_dense_tc_rescore_append(cand, cnt, kGmCap, emb, ld, q_dev, n_real, k,
                                        emb_norm_max * eps_rel, thr_key, out, flags, stream,
                                        n_flagged ? gate_ctr + 8 : nullptr, n_flagged, flagged);
}

}  // namespace anr
