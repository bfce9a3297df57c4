// Shared device helpers: sortable keys, warp primitives, mbarrier / bulk-copy PTX.
// sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace anr {

constexpr unsigned kFullMask = 0xffffffffu;

// ---- sortable (score, id) keys ---------------------------------------------
// key = orderable(score) << 32 | (0xFFFFFFFF - id): a larger key is a better hit
// (higher score; equal scores -> lower id first).  Key 0 is the empty slot: no
// finite/infinite score maps to orderable 0.
__host__ __device__ __forceinline__ uint32_t f32_to_ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
  score = score + 0.0f;  // -0.0 -> +0.0 so both zeros tie, as they do for numpy
  return (static_cast<uint64_t>(f32_to_ord(score)) << 32) | static_cast<uint64_t>(0xffffffffu - id);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
  return ord_to_f32(static_cast<uint32_t>(key >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t key) {
  return 0xffffffffu - static_cast<uint32_t>(key & 0xffffffffu);
}

// ---- warp primitives ---------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t other = __shfl_xor_sync(kFullMask, v, o);
    v = other < v ? other : v;
  }
  return v;
}

// ---- shared-memory addresses, mbarrier, 1-D bulk async copy (TMA engine) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- programmatic dependent launch (chains of short kernels on one stream) ----
// A step is ~10 dependent kernels per chain, most of them a few microseconds long: what separates
// them (drain, flush, launch, CTA scheduling) costs as much as they do.  Kernels launched through
// launch_chain() may become resident while their predecessor still runs; pdl_wait() -- the FIRST
// statement of every such kernel -- then blocks until the predecessor grid has completed and its
// writes are visible, so the stream order of all memory effects is unchanged.  pdl_trigger()
// lets the successor be scheduled (it fires once every CTA of the grid has called it or exited);
// placed after pdl_wait(), at most one successor waits on the SMs at a time.  Long kernels call
// it at their end, so that nothing parks beside them.  Without the launch attribute both are
// no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float4 lds128(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace anr
