// Top-k building blocks on 64-bit sortable keys (anr_common.cuh):
//   * a per-warp unsorted k-list in shared memory with a running threshold
//     (almost every row is rejected by one register compare);
//   * a block-wide bitonic sort (descending) in shared memory;
//   * the final per-query selection kernel over a candidate array.
#pragma once
#include "anr_common.cuh"
#include "anr_internal.h"

namespace anr {

constexpr int kMaxFusedK = 128;      // fused (threshold-list) top-k handles k <= this
constexpr int kFinalThreads = 1024;  // topk_final_kernel block size
constexpr int kFinalSortCap = 4096;  // keys the final kernel sorts in shared memory

// One result entry (key 0 = empty slot) into whichever output arrays the caller asked for.
__device__ __forceinline__ void emit_entry(uint64_t key, int64_t slot, const TopkOut& o) {
  const bool valid = key != 0ull;
  uint32_t id = key_id(key);
  if (valid)
    id = o.id_map ? static_cast<uint32_t>(o.id_map[id]) : static_cast<uint32_t>(id + o.id_base);
  if (o.keys) o.keys[slot] = valid ? ((key & 0xffffffff00000000ull) | (0xffffffffu - id)) : 0ull;
  if (o.scores) o.scores[slot] = valid ? key_score(key) : 0.f;
  if (o.ids) o.ids[slot] = valid ? static_cast<int32_t>(id) : -1;
}

// Replace the minimum of list[0..k) by `key` (caller guarantees key > current
// minimum or the list has empty slots) and return the new minimum = the warp's
// new acceptance threshold.  All 32 lanes call with the same arguments.
__device__ __forceinline__ uint64_t warp_list_insert(uint64_t* list, int k, uint64_t key, int lane) {
  uint64_t mn = ~0ull;
  int mp = 0x7fffffff;
  for (int i = lane; i < k; i += 32) {
    uint64_t v = list[i];
    if (v < mn) { mn = v; mp = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t omn = __shfl_xor_sync(kFullMask, mn, o);
    int omp = __shfl_xor_sync(kFullMask, mp, o);
    if (omn < mn || (omn == mn && omp < mp)) { mn = omn; mp = omp; }
  }
  if (lane == 0) list[mp] = key;
  __syncwarp();
  uint64_t nm = ~0ull;
  for (int i = lane; i < k; i += 32) {
    uint64_t v = list[i];
    nm = v < nm ? v : nm;
  }
  return warp_min_u64(nm);
}

// Out-of-line copy for hot loops: keeps the rarely taken insert path out of the instruction
// stream of the scan (32 inlined copies made the dense scan I-cache bound, see profiles/).
static __device__ __noinline__ uint64_t warp_list_insert_cold(uint64_t* list, int k, uint64_t key,
                                                       int lane) {
  if (k <= 32) {  // one lane per slot: ~5x fewer instructions than the strided loops
    uint64_t v = lane < k ? list[lane] : ~0ull;
    uint64_t mn = v;
    int ml = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t om = __shfl_xor_sync(kFullMask, mn, o);
      const int ol = __shfl_xor_sync(kFullMask, ml, o);
      if (om < mn || (om == mn && ol < ml)) { mn = om; ml = ol; }
    }
    if (lane == ml) { list[lane] = key; v = key; }
    __syncwarp();
    return warp_min_u64(v);
  }
  return warp_list_insert(list, k, key, lane);
}

// V partial sums per lane -> one warp total per lane: after log2(V) exchange stages lane l
// holds the total of element l / (32 / V) (V - 1 shuffles instead of 5 * V), the remaining
// stages are a plain butterfly inside each group of 32 / V lanes.
template <int V>
__device__ __forceinline__ float warp_transpose_reduce(float* v, int lane) {
  static_assert(V == 1 || V == 2 || V == 4 || V == 8 || V == 16 || V == 32, "V must divide 32");
  int o = 16;
#pragma unroll
  for (int c = V; c > 1; c >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < c / 2; ++j) {
      const float send = upper ? v[j] : v[j + c / 2];
      const float keep = upper ? v[j + c / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(kFullMask, send, o);
    }
    o >>= 1;
  }
  float total = v[0];
  for (; o > 0; o >>= 1) total += __shfl_xor_sync(kFullMask, total, o);
  return total;
}

// Lanes hold DIFFERENT candidate keys: insert every lane's key that beats thr.
__device__ __forceinline__ void warp_list_offer(uint64_t* list, int k, uint64_t my_key,
                                                uint64_t& thr, int lane) {
  unsigned pending = __ballot_sync(kFullMask, my_key > thr);
  while (pending) {
    int src = __ffs(pending) - 1;
    pending &= pending - 1;
    uint64_t cand = __shfl_sync(kFullMask, my_key, src);
    if (cand > thr) thr = warp_list_insert(list, k, cand, lane);
  }
}

// Bitonic sort of keys[0..n_pow2) descending; every thread of the block calls.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* keys, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        int lo = ((t / stride) * (stride << 1)) + (t % stride);
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// k-th largest of the per-thread best keys of a block of NWARPS warps (0 when fewer than k
// threads hold a key): with every thread owning a disjoint set of candidates, k distinct
// candidates reach it, so it bounds the k-th best candidate from below.  Every thread calls;
// tbest holds NWARPS * 32 keys of scratch, *s_out one key.
//   k <= 32: every warp sorts its 32 keys in registers (shuffles, no barrier), warp 0 then pops
//            the largest run head k times; larger k: block-wide bitonic sort (measured: a pop is
//            a ~175-cycle dependent chain, so beyond a few dozen pops the 45 stages of a 512-key
//            bitonic sort win: rank 155 cost +26 us per dense pass with pops).
template <int NWARPS>
__device__ __forceinline__ uint64_t block_kth_of_thread_bests(uint64_t best, int k, uint64_t* tbest,
                                                              uint64_t* s_out) {
  static_assert(NWARPS <= 32, "one lane of warp 0 per sorted run");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k <= 32) {
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
      int head = 0;   // lanes 0..NWARPS-1: read position in warp `lane`'s sorted run
      uint64_t kth = 0ull;
      for (int it = 0; it < k; ++it) {
        const uint64_t c = (lane < NWARPS && head < 32) ? tbest[lane * 32 + head] : 0ull;
        uint64_t m = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint64_t other = __shfl_xor_sync(kFullMask, m, o);
          m = other > m ? other : m;
        }
        kth = m;
        if (m == 0ull) break;               // fewer than k threads hold a key
        // equal keys in two runs (possible only for keys without unique ids): one lane advances
        const unsigned who = __ballot_sync(kFullMask, c == m);
        if (lane == __ffs(who) - 1) ++head;
      }
      if (lane == 0) *s_out = kth;
    }
    __syncthreads();
  } else {
    tbest[threadIdx.x] = best;
    block_bitonic_sort_desc(tbest, NWARPS * 32);
    if (threadIdx.x == 0) *s_out = k <= NWARPS * 32 ? tbest[k - 1] : 0ull;
    __syncthreads();
  }
  return *s_out;
}

}  // namespace anr
