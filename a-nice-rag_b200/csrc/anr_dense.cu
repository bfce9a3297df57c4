// Dense inner-product scan fused with top-k (HBM-bound path, 1..8 queries per pass).
//
// Replaces np.stack + np.dot + argpartition of the reference
// (src/search_engine.py:80-87, :128-135).  One persistent CTA per SM:
//   * a producer lane streams contiguous groups of RW rows HBM -> shared memory with
//     1-D bulk async copies (TMA engine, cp.async.bulk + mbarrier complete_tx) into a
//     ring of up to 32 stages, so the bytes in flight per SM (~200 KB) are set by the
//     ring and not by register pressure;
//   * 8 consumer warps each own every 8th stage (a stage is always refilled for the same
//     warp, so the mbarrier parity protocol can never be lapped): a warp takes its RW rows, every lane
//     reads 128-bit words of the rows and of the staged queries (conflict-free
//     LDS.128; the query words are reused by RW rows), fp32 FMA accumulation,
//     xor-butterfly warp reduction;
//   * every finished (row, query) score is packed into a sortable key and compared
//     with the warp's running k-th best (one register compare rejects almost every
//     row); survivors go into a per-warp k-list in shared memory;
//   * at the end the CTA bitonic-sorts its 8 lists per query and writes k
//     candidates per query; topk_final_kernel (anr_topk.cu) merges the CTAs.
// Algorithmic HBM bytes: n * ld * 4 per pass (each row read exactly once).
#include <cstdlib>

#include "anr_internal.h"
#include "anr_topk.cuh"

namespace anr {

constexpr int kScanConsumerWarps = 8;
constexpr int kScanThreads = (kScanConsumerWarps + 1) * 32;
constexpr int kScanMaxStages = 32;

struct ScanLayout {
  int q_off, ring_off, list_off, bar_off, total_bytes;
  int stage_floats, n_stages, list_cap;
  int n_cwarps;  // consumer warps that own stages (n_stages is a multiple of it)
};

// One full pass over the corpus by this CTA (see the file header).  Every thread of the CTA
// calls it; shared memory is (re)initialised on entry, so it can be called repeatedly.
template <int NQ, int RW, bool EMIT_ALL>
__device__ __forceinline__ void dense_scan_body(const float* __restrict__ emb, int64_t n, int ld,
                                                const float* __restrict__ q, int k,
                                                const uint32_t* __restrict__ mask,
                                                uint64_t* __restrict__ out, int64_t out_stride_q,
                                                const ScanLayout& L, unsigned char* smem) {
  float* qs = reinterpret_cast<float*>(smem + L.q_off);
  float* ring = reinterpret_cast<float*>(smem + L.ring_off);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + L.list_off);  // [NQ][list_cap]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty = full + kScanMaxStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (n + RW - 1) / RW;  // one tile = RW consecutive rows = one stage
  // tiles of this CTA: blockIdx.x + it * gridDim.x, it = 0 .. my_tiles-1
  const int64_t my_tiles =
      n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  for (int i = threadIdx.x; i < NQ * ld; i += blockDim.x) qs[i] = q[i];
  if (!EMIT_ALL)
    for (int i = threadIdx.x; i < NQ * L.list_cap; i += blockDim.x) lists[i] = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < L.n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kScanConsumerWarps) {
    // ---- producer: one lane feeds the ring ----
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        mbar_wait(&empty[s], ph ^ 1u);
        const int64_t row0 = (blockIdx.x + it * gridDim.x) * RW;
        const int64_t left = n - row0;
        const int rows = left < RW ? static_cast<int>(left) : RW;
        const uint32_t bytes = static_cast<uint32_t>(rows) * ld * 4u;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(ring + static_cast<size_t>(s) * L.stage_floats, emb + row0 * ld, bytes, &full[s]);
        if (++s == L.n_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ---- consumers: warp w owns local tiles it = w, w + n_cwarps, ... ----
    uint64_t thr = 0;  // this warp's k-th best key so far for THIS LANE's query (see below)
    // stage of local tile it is it % n_stages; n_stages is a multiple of n_cwarps, so warp w
    // only ever sees stages w, w + n_cwarps, ... and the parity flips when that walk wraps
    int s = warp - L.n_cwarps;
    uint32_t ph = 0;
    for (int64_t it = warp; warp < L.n_cwarps && it < my_tiles; it += L.n_cwarps) {
      s += L.n_cwarps;
      if (s >= L.n_stages) { s -= L.n_stages; ph ^= 1u; }
      const float* tile = ring + static_cast<size_t>(s) * L.stage_floats;
      const int64_t row0 = (blockIdx.x + it * gridDim.x) * RW;
      const int64_t left = n - row0;
      const int rows = left < RW ? static_cast<int>(left) : RW;
      mbar_wait(&full[s], ph);

      float acc[RW][NQ];
#pragma unroll
      for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) acc[r][qi] = 0.f;
      // rows past the end of a short last tile alias the last valid row (discarded below)
      const float* rp[RW];
#pragma unroll
      for (int r = 0; r < RW; ++r) rp[r] = tile + static_cast<size_t>(min(r, rows - 1)) * ld;

#pragma unroll 2
      for (int c = lane * 4; c < ld; c += 128) {
        float4 e[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) e[r] = lds128(rp[r] + c);
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          const float4 qv = lds128(qs + qi * ld + c);
#pragma unroll
          for (int r = 0; r < RW; ++r) {
            acc[r][qi] = fmaf(e[r].x, qv.x, acc[r][qi]);
            acc[r][qi] = fmaf(e[r].y, qv.y, acc[r][qi]);
            acc[r][qi] = fmaf(e[r].z, qv.z, acc[r][qi]);
            acc[r][qi] = fmaf(e[r].w, qv.w, acc[r][qi]);
          }
        }
      }
      // the stage's data now lives in registers: hand the slot back before reducing
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);

      // V = RW * NQ partial sums per lane -> ONE finished score per lane: lane l ends up with
      // the warp total of (row r, query qi) = divmod(l / G, NQ), replicated over G = 32 / V lanes
      constexpr int V = RW * NQ;
      constexpr int G = 32 / V;
      const float total = warp_transpose_reduce<V>(&acc[0][0], lane);
      const int idx = lane / G;
      const int r = idx / NQ, qi = idx % NQ;
      const int64_t gi = row0 + r;
      const bool leader = (lane % G) == 0 && r < rows;
      if (EMIT_ALL) {
        if (leader) {
          bool ok = true;
          if (mask) ok = (__ldg(mask + (gi >> 5)) >> (gi & 31)) & 1u;
          out[qi * out_stride_q + gi] = ok ? make_key(total, static_cast<uint32_t>(gi)) : 0ull;
        }
      } else {
        const uint64_t key = leader ? make_key(total, static_cast<uint32_t>(gi)) : 0ull;
        unsigned pending = __ballot_sync(kFullMask, key > thr);
        while (pending) {  // rare: almost every score is rejected by the compare above
          const int src = __ffs(pending) - 1;
          pending &= pending - 1;
          const uint64_t cand = __shfl_sync(kFullMask, key, src);
          const uint64_t cur = __shfl_sync(kFullMask, thr, src);
          if (cand <= cur) continue;  // an earlier insert of this pass raised the bar
          const uint32_t row = key_id(cand);
          if (mask && !((__ldg(mask + (row >> 5)) >> (row & 31)) & 1u)) continue;
          const int sq = (src / G) % NQ;
          const uint64_t nt =
              warp_list_insert_cold(lists + sq * L.list_cap + warp * k, k, cand, lane);
          if (qi == sq) thr = nt;
        }
      }
    }
  }

  if (EMIT_ALL) return;
  __syncthreads();
  // ---- CTA merge: sort the 8 per-warp lists of each query, emit the best k ----
  for (int qi = 0; qi < NQ; ++qi) {
    uint64_t* lq = lists + qi * L.list_cap;
    block_bitonic_sort_desc(lq, L.list_cap);
    for (int i = threadIdx.x; i < k; i += blockDim.x)
      out[qi * out_stride_q + static_cast<int64_t>(blockIdx.x) * k + i] = lq[i];
    __syncthreads();
  }
}

template <int NQ, int RW, bool EMIT_ALL>
__global__ void __launch_bounds__(kScanThreads, 1)
dense_scan_kernel(const float* __restrict__ emb, int64_t n, int ld, const float* __restrict__ q,
                  int k, const uint32_t* __restrict__ mask, uint64_t* __restrict__ out,
                  int64_t out_stride_q, ScanLayout L) {
  extern __shared__ __align__(128) unsigned char smem[];
  dense_scan_body<NQ, RW, EMIT_ALL>(emb, n, ld, q, k, mask, out, out_stride_q, L, smem);
}

// Exact rescan of the queries the tensor-core path could not certify (anr_dense_tc.cu): the
// list of flagged queries lives on the device, so no host round trip is needed -- one launch,
// which returns at once when the list is empty (the common case) and otherwise runs one full
// single-query pass per flagged query.  cand: [flagged slot][cta * k + i].
template <int RW>
__global__ void __launch_bounds__(kScanThreads, 1)
dense_scan_flagged_kernel(const float* __restrict__ emb, int64_t n, int ld,
                          const float* __restrict__ q_all, const int32_t* __restrict__ n_flagged,
                          const int32_t* __restrict__ flagged, int k,
                          const uint32_t* __restrict__ mask, uint64_t* __restrict__ cand,
                          int64_t cand_stride, ScanLayout L) {
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_wait();
  pdl_trigger();
  const int nf = *n_flagged;
  for (int f = 0; f < nf; ++f) {
    dense_scan_body<1, RW, false>(emb, n, ld, q_all + static_cast<size_t>(flagged[f]) * ld, k, mask,
                                  cand + f * cand_stride, cand_stride, L, smem);
    __syncthreads();
    if (threadIdx.x == 0) {   // the next pass initialises the barriers again
      uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
      for (int s = 0; s < 2 * kScanMaxStages; ++s)
        if (s < L.n_stages || (s >= kScanMaxStages && s < kScanMaxStages + L.n_stages))
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[s])) : "memory");
    }
    __syncthreads();
  }
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

static bool make_scan_layout(const DeviceProps& dp, int ld, int nq, int rw, int k, bool emit_all,
                             ScanLayout* L) {
  L->q_off = 0;
  L->ring_off = align_up(nq * ld * 4, 128);
  L->list_cap = emit_all ? 0 : next_pow2(kScanConsumerWarps * k);
  const int list_bytes = nq * L->list_cap * 8;
  const int bar_bytes = 2 * kScanMaxStages * 8;
  const int64_t stage_bytes = static_cast<int64_t>(rw) * ld * 4;
  // mbarrier transaction counts are 20-bit: one stage must stay below 1 MiB
  if (stage_bytes >= (1 << 20)) return false;
  const int64_t avail =
      static_cast<int64_t>(dp.max_smem_optin) - L->ring_off - list_bytes - bar_bytes - 512;
  int64_t n_stages = avail / stage_bytes;
  if (n_stages < 3) return false;
  if (n_stages > kScanMaxStages) n_stages = kScanMaxStages;
  // stage s always serves consumer warp s % n_cwarps
  L->n_cwarps = n_stages < kScanConsumerWarps ? static_cast<int>(n_stages) : kScanConsumerWarps;
  n_stages = n_stages / L->n_cwarps * L->n_cwarps;
  L->stage_floats = rw * ld;
  L->n_stages = static_cast<int>(n_stages);
  L->list_off = align_up(L->ring_off + static_cast<int>(n_stages * stage_bytes), 16);
  L->bar_off = align_up(L->list_off + list_bytes, 16);
  L->total_bytes = L->bar_off + bar_bytes;
  return L->total_bytes <= dp.max_smem_optin;
}

// Rows per stage.
static int choose_rw(const DeviceProps& dp, int ld, int nq, int k, bool emit_all) {
  if (const char* force = getenv("ANR_SCAN_RW")) {  // tuning knob for profiling runs
    const int rw = atoi(force);
    ScanLayout L;
    if ((rw == 1 || rw == 2 || rw == 4) && make_scan_layout(dp, ld, nq, rw, k, emit_all, &L))
      return rw;
  }
  // measured on B200 (profiles/): 16 KB bulk copies (RW = 4 at D = 1024) with 8 stages reach
  // 7.1 TB/s, 8 KB copies with 24 stages 5.9 TB/s, 4 KB copies 3.0 TB/s -> largest RW that
  // still leaves one stage per consumer warp
  int fallback = 0;
  for (int rw = 4; rw >= 1; rw >>= 1) {
    ScanLayout L;
    if (!make_scan_layout(dp, ld, nq, rw, k, emit_all, &L)) continue;
    if (L.n_stages >= kScanConsumerWarps) return rw;
    if (!fallback) fallback = rw;
  }
  return fallback;
}

int dense_scan_max_grid(const DeviceProps& dp) { return dp.sm_count; }

int dense_scan_max_queries(const DeviceProps& dp, int ld, int k, bool emit_all) {
  ScanLayout L;
  for (int nq = 8; nq >= 1; nq >>= 1)
    if (make_scan_layout(dp, ld, nq, 1, k, emit_all, &L)) return nq;
  return 0;
}

template <int NQ, int RW, bool EMIT_ALL>
static cudaError_t launch_scan_t(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                 const float* q_dev, int k, const uint32_t* mask, uint64_t* out,
                                 int64_t out_stride_q, int* grid_out, cudaStream_t stream) {
  ScanLayout L;
  if (!make_scan_layout(dp, ld, NQ, RW, k, EMIT_ALL, &L)) return cudaErrorInvalidConfiguration;
  auto kern = dense_scan_kernel<NQ, RW, EMIT_ALL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       L.total_bytes);
  if (e != cudaSuccess) return e;
  const int64_t n_tiles = (n + RW - 1) / RW;
  int grid = static_cast<int>(n_tiles < dp.sm_count ? n_tiles : dp.sm_count);
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = grid;
  kern<<<grid, kScanThreads, L.total_bytes, stream>>>(emb, n, ld, q_dev, k, mask, out,
                                                      out_stride_q, L);
  return cudaGetLastError();
}

template <int NQ, bool EMIT_ALL>
static cudaError_t dispatch_rw(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                               const float* q_dev, int k, const uint32_t* mask, uint64_t* out,
                               int64_t out_stride_q, int* grid_out, cudaStream_t stream) {
  switch (choose_rw(dp, ld, NQ, k, EMIT_ALL)) {
    case 4: return launch_scan_t<NQ, 4, EMIT_ALL>(dp, emb, n, ld, q_dev, k, mask, out, out_stride_q, grid_out, stream);
    case 2: return launch_scan_t<NQ, 2, EMIT_ALL>(dp, emb, n, ld, q_dev, k, mask, out, out_stride_q, grid_out, stream);
    case 1: return launch_scan_t<NQ, 1, EMIT_ALL>(dp, emb, n, ld, q_dev, k, mask, out, out_stride_q, grid_out, stream);
    default: return cudaErrorInvalidConfiguration;
  }
}

template <bool EMIT_ALL>
static cudaError_t dispatch_scan(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                 const float* q_dev, int nq, int k, const uint32_t* mask,
                                 uint64_t* out, int64_t out_stride_q, int* grid_out,
                                 cudaStream_t stream) {
  switch (nq) {
    case 1: return dispatch_rw<1, EMIT_ALL>(dp, emb, n, ld, q_dev, k, mask, out, out_stride_q, grid_out, stream);
    case 2: return dispatch_rw<2, EMIT_ALL>(dp, emb, n, ld, q_dev, k, mask, out, out_stride_q, grid_out, stream);
    case 4: return dispatch_rw<4, EMIT_ALL>(dp, emb, n, ld, q_dev, k, mask, out, out_stride_q, grid_out, stream);
    case 8: return dispatch_rw<8, EMIT_ALL>(dp, emb, n, ld, q_dev, k, mask, out, out_stride_q, grid_out, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_dense_scan_topk(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                   const float* q_dev, int nq, int k, const uint32_t* mask,
                                   uint64_t* cand, int64_t cand_stride_q, int* grid_out,
                                   cudaStream_t stream) {
  if (k < 1 || k > kMaxFusedK) return cudaErrorInvalidValue;
  return dispatch_scan<false>(dp, emb, n, ld, q_dev, nq, k, mask, cand, cand_stride_q, grid_out,
                              stream);
}

template <int RW>
static cudaError_t launch_flagged_t(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                    const float* q_all, const int32_t* n_flagged,
                                    const int32_t* flagged, int k, const uint32_t* mask,
                                    uint64_t* cand, int64_t cand_stride, cudaStream_t stream) {
  ScanLayout L;
  if (!make_scan_layout(dp, ld, 1, RW, k, false, &L)) return cudaErrorInvalidConfiguration;
  auto kern = dense_scan_flagged_kernel<RW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       L.total_bytes);
  if (e != cudaSuccess) return e;
  const int64_t n_tiles = (n + RW - 1) / RW;
  int grid = static_cast<int>(n_tiles < dp.sm_count ? n_tiles : dp.sm_count);
  if (grid < 1) grid = 1;
  return launch_chain(kern, dim3(grid), dim3(kScanThreads), static_cast<size_t>(L.total_bytes), stream,
                      emb, n, ld, q_all, n_flagged, flagged, k, mask, cand, cand_stride, L);
}

// grid the flagged rescan uses = candidates per flagged query / k
int dense_scan_flagged_grid(const DeviceProps& dp, int64_t n, int ld, int k) {
  const int rw = choose_rw(dp, ld, 1, k, false);
  if (rw < 1) return 0;
  const int64_t n_tiles = (n + rw - 1) / rw;
  return static_cast<int>(n_tiles < dp.sm_count ? (n_tiles < 1 ? 1 : n_tiles) : dp.sm_count);
}

cudaError_t launch_dense_scan_flagged(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                      const float* q_all, const int32_t* n_flagged,
                                      const int32_t* flagged, int k, const uint32_t* mask,
                                      uint64_t* cand, int64_t cand_stride, cudaStream_t stream) {
  if (k < 1 || k > kMaxFusedK) return cudaErrorInvalidValue;
  switch (choose_rw(dp, ld, 1, k, false)) {
    case 4: return launch_flagged_t<4>(dp, emb, n, ld, q_all, n_flagged, flagged, k, mask, cand, cand_stride, stream);
    case 2: return launch_flagged_t<2>(dp, emb, n, ld, q_all, n_flagged, flagged, k, mask, cand, cand_stride, stream);
    case 1: return launch_flagged_t<1>(dp, emb, n, ld, q_all, n_flagged, flagged, k, mask, cand, cand_stride, stream);
    default: return cudaErrorInvalidConfiguration;
  }
}

cudaError_t launch_dense_scan_all(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                  const float* q_dev, int nq, const uint32_t* mask, uint64_t* keys,
                                  int64_t keys_stride_q, cudaStream_t stream) {
  return dispatch_scan<true>(dp, emb, n, ld, q_dev, nq, 1, mask, keys, keys_stride_q, nullptr,
                             stream);
}

}  // namespace anr
