// C ABI (include/anr_b200.h): contexts, index objects, argument staging and the
// kernel pipelines of each entry point.  No C++ exception crosses the boundary.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/anr_b200.h"
#include "anr_internal.h"
#include "anr_topk.cuh"

using namespace anr;

// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* what, const char* detail = nullptr) {
  g_last_error = what;
  if (detail) {
    g_last_error += ": ";
    g_last_error += detail;
  }
  return code;
}
static int fail_cuda(const char* what, cudaError_t e) {
  // leave the sticky error state readable but do not keep it queued for the next call
  cudaGetLastError();
  return fail(e == cudaErrorMemoryAllocation ? ANR_ERR_OOM : ANR_ERR_CUDA, what,
              cudaGetErrorString(e));
}
namespace anr {
int set_error(int code, const char* what, const char* detail) { return fail(code, what, detail); }
}  // namespace anr
#define ANR_CUDA(expr)                                       \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return fail_cuda(#expr, _e);      \
  } while (0)

// ---------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------
struct EventPair {
  cudaEvent_t start = nullptr, stop = nullptr;
};
struct anr_ctx {
  DeviceProps dp;
  cudaStream_t stream = nullptr;
  // BM25 of a hybrid query runs on `side` under the dense scan (fork/join through the events)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mid = nullptr;
  unsigned char* ws = nullptr;  // device scratch, grown on demand
  size_t ws_bytes = 0;
  // profiling (anr_ctx_profile_*): event pairs recorded around the dominant kernels
  bool profiling = false;
  // anr_ctx_set_beside_dense: this context's BM25 searches run next to a dense pass issued
  // through another context / stream (the sharded search does that)
  bool beside_dense = false;
  // kind 0 = dense scan kernel, 1 = BM25 score kernel, 2 = a whole tensor-core pass
  // (sample pre-pass + threshold + scan + rescoring)
  std::vector<EventPair> pool[3];  // created lazily, reused after every read
  size_t used[3] = {0, 0, 0};
  // The scratch arena is shared by every call on this context, and calls are asynchronous on the
  // CALLER's stream: a call issued on another stream than the previous one (torch's current
  // stream, then NULL = the context's own non-blocking stream) would otherwise run beside it in
  // the same scratch.  Every entry point waits on ev_last when the stream changes (StreamOrder).
  cudaEvent_t ev_last = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_valid = false;
  // anr_ctx_last_rerun: device counters of the last call's rerun lists (in the scratch buffer: valid
  // until the next call on this context), null when that call had no such list
  const int32_t* last_dense_flagged = nullptr;
  const int32_t* last_bm25_flagged = nullptr;
  // anr_ctx_timeline_*: where the parts of ONE hybrid step start and end (see the header)
  bool timeline = false;
  cudaEvent_t tl[ANR_TIMELINE_MARKS] = {};
  bool tl_set[ANR_TIMELINE_MARKS] = {};
};

struct anr_dense {
  int device = 0;
  float* emb = nullptr;
  int64_t n = 0;
  int32_t d = 0;
  int32_t ld = 0;  // leading dimension in floats (d rounded up to a multiple of 4)
  bool owned = true;
  // max row norm, for the tf32 error bound of the tensor-core scan (computed on first need)
  mutable float norm_max = 0.f;
  mutable bool norm_valid = false;
  // bf16 copy of emb ([n, ld]) for the GEMM path, built on first use when enabled
  // (anr_dense_set_shadow / ANR_TC_BF16=1)
  mutable void* shadow = nullptr;
  mutable bool want_shadow = false;
  // indices are shared between contexts / threads: the lazily built members above are guarded
  mutable std::mutex lazy;
};

struct anr_bm25 {
  int device = 0;
  int64_t* term_ptr = nullptr;
  int32_t* post_doc = nullptr;
  float* post_w = nullptr;
  float* idf = nullptr;
  int32_t n_terms = 0;
  int32_t n_docs = 0;
  int64_t nnz = 0;
  // dense rows of the head terms for the pruned top-k scan (anr_bm25.cu), built on first use
  mutable uint8_t* head_slot = nullptr;    // [n_terms]
  mutable int32_t* head_terms = nullptr;   // [n_head]
  mutable float* head_w = nullptr;         // [n_head][head_ld]
  mutable float* head_max = nullptr;       // [n_head]
  mutable int64_t head_ld = 0;
  mutable int32_t n_head = 0;
  mutable bool head_built = false;         // slot map + allocation exist
  mutable bool head_filled = false;        // rows hold the current posting weights
  // candidate-driven top-k (anr_bm25_ms.cu): largest posting weight per term, built on first use
  mutable float* term_maxw = nullptr;      // [n_terms] (+ one int behind it: the negative-idf probe)
  mutable int64_t* bkt_off = nullptr;      // [n_terms]
  mutable uint8_t* bkt_shift = nullptr;    // [n_terms]
  mutable int32_t* bkt = nullptr;          // bucket tables
  mutable bool ms_ready = false;
  mutable bool neg_idf = false;            // some idf < 0: the bounds do not hold, tiled scan only
  mutable std::mutex lazy;                 // guards the lazily built members (shared index)
};

namespace {

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Bump allocator over the context's scratch buffer.
struct Arena {
  unsigned char* base;
  size_t cap;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
};
inline size_t padded(size_t bytes) { return (bytes + 255) & ~static_cast<size_t>(255); }

int ws_reserve(anr_ctx* ctx, size_t bytes) {
  ctx->last_dense_flagged = nullptr;   // (a new call: the previous call's rerun counters are gone)
  ctx->last_bm25_flagged = nullptr;
  bytes += 4096;
  if (bytes <= ctx->ws_bytes) return ANR_OK;
  // earlier work (on the context's stream or a caller's) may still be using the old buffer
  ANR_CUDA(cudaDeviceSynchronize());
  if (ctx->ws) cudaFree(ctx->ws);
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  size_t want = std::max(bytes, static_cast<size_t>(8) << 20);
  want += want / 4;
  ANR_CUDA(cudaMalloc(&ctx->ws, want));
  ctx->ws_bytes = want;
  return ANR_OK;
}

// Brackets one kernel launch with events when profiling is on.
struct ProfileScope {
  anr_ctx* ctx;
  int kind;
  cudaStream_t stream;
  EventPair* ev = nullptr;
  ProfileScope(anr_ctx* c, int k, cudaStream_t s) : ctx(c), kind(k), stream(s) {
    if (!ctx->profiling) return;
    if (ctx->used[kind] == ctx->pool[kind].size()) {
      if (ctx->pool[kind].size() >= 65536) return;  // stop recording, keep running
      EventPair p;
      if (cudaEventCreate(&p.start) != cudaSuccess || cudaEventCreate(&p.stop) != cudaSuccess) {
        cudaGetLastError();
        return;
      }
      ctx->pool[kind].push_back(p);
    }
    ev = &ctx->pool[kind][ctx->used[kind]++];
    cudaEventRecord(ev->start, stream);
  }
  ~ProfileScope() {
    if (ev) cudaEventRecord(ev->stop, stream);
  }
};

// An event pair the callee records itself (around one kernel inside a multi-kernel launcher).
EventPair* profile_take(anr_ctx* ctx, int kind) {
  if (!ctx->profiling) return nullptr;
  if (ctx->used[kind] == ctx->pool[kind].size()) {
    if (ctx->pool[kind].size() >= 65536) return nullptr;
    EventPair p;
    if (cudaEventCreate(&p.start) != cudaSuccess || cudaEventCreate(&p.stop) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    ctx->pool[kind].push_back(p);
  }
  return &ctx->pool[kind][ctx->used[kind]++];
}

// Timeline mark `id` (include/anr_b200.h) on `stream`; a no-op unless anr_ctx_timeline_enable is on.
void tl_mark(anr_ctx* ctx, int id, cudaStream_t stream) {
  if (!ctx->timeline || id < 0 || id >= ANR_TIMELINE_MARKS) return;
  if (!ctx->tl[id] && cudaEventCreate(&ctx->tl[id]) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  if (cudaEventRecord(ctx->tl[id], stream) == cudaSuccess) ctx->tl_set[id] = true;
  else cudaGetLastError();
}

// Orders a call after the previous call of the same context when that one was enqueued on a
// different stream (same stream: stream order already does it).  Nothing is recorded while the
// stream is being captured into a graph: a captured step owns its context and its stream.
struct StreamOrder {
  anr_ctx* ctx;
  cudaStream_t stream;
  bool live = false;
  StreamOrder(anr_ctx* c, cudaStream_t s) : ctx(c), stream(s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return; }
    if (st != cudaStreamCaptureStatusNone) return;
    live = true;
    if (ctx->last_valid && ctx->last_stream != stream && ctx->ev_last)
      cudaStreamWaitEvent(stream, ctx->ev_last, 0);
  }
  ~StreamOrder() {
    if (!live || !ctx->ev_last) return;
    if (cudaEventRecord(ctx->ev_last, stream) == cudaSuccess) {
      ctx->last_stream = stream;
      ctx->last_valid = true;
    } else {
      cudaGetLastError();
    }
  }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Queries are scanned in groups of gmax (1, 2, 4 or 8); a short batch is one group of the next
// power of two, a long one is padded to a multiple of gmax (pad rows are zero vectors).
inline int pad_queries(int nq, int gmax) {
  if (nq <= gmax) {
    int p = 1;
    while (p < nq) p <<= 1;
    return p;
  }
  return (nq + gmax - 1) / gmax * gmax;
}
inline int dense_group(const anr_ctx* ctx, const anr_dense* ix, int k) {
  const bool all = k > kMaxFusedK;
  return dense_scan_max_queries(ctx->dp, ix->ld, all ? 1 : k, all);
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}
__global__ void set_f64_pair_kernel(double* out, double a, double b) {
  out[0] = a;
  out[1] = b;
}
__global__ void set_i32_pair_kernel(int32_t* out, int32_t a, int32_t b) {
  out[0] = a;
  out[1] = b;
}
// BM25 results carry document ids in their own space: translate them to row ids of the
// dense index' space before fusion (done by TopkOut::id_map), nothing else needed here.

// Copy a [rows, d] fp32 matrix (host or device) into a [rows, ld] device matrix.
cudaError_t copy_rows(float* dst, int ld, const float* src, int d, int64_t rows,
                      cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (ld == d)
    return cudaMemcpyAsync(dst, src, static_cast<size_t>(rows) * d * 4, cudaMemcpyDefault, stream);
  cudaError_t e = cudaMemsetAsync(dst, 0, static_cast<size_t>(rows) * ld * 4, stream);
  if (e != cudaSuccess) return e;
  return cudaMemcpy2DAsync(dst, static_cast<size_t>(ld) * 4, src, static_cast<size_t>(d) * 4,
                           static_cast<size_t>(d) * 4, static_cast<size_t>(rows),
                           cudaMemcpyDefault, stream);
}

// Host-or-device output: device pointers are written in place; host pointers get a
// scratch twin that is copied back (and the call synchronises) at the end.
template <typename T>
struct OutBuf {
  T* user = nullptr;
  T* dev = nullptr;
  size_t count = 0;
  bool staged = false;
};
template <typename T>
size_t out_need(const T* user, size_t count) {
  return (user && !is_device_ptr(user)) ? padded(count * sizeof(T)) + 256 : 0;
}
template <typename T>
OutBuf<T> out_make(Arena& a, T* user, size_t count) {
  OutBuf<T> o;
  o.user = user;
  o.count = count;
  if (!user) return o;
  if (is_device_ptr(user)) {
    o.dev = user;
  } else {
    o.dev = a.take<T>(count);
    o.staged = true;
  }
  return o;
}
template <typename T>
cudaError_t out_flush(const OutBuf<T>& o, cudaStream_t stream, bool* any_host) {
  if (!o.staged) return cudaSuccess;
  *any_host = true;
  return cudaMemcpyAsync(o.user, o.dev, o.count * sizeof(T), cudaMemcpyDeviceToHost, stream);
}

// ---- dense top-k pipeline on device buffers -----------------------------------------
// q_dev: [pad_queries(nq), ld].  Writes results through `out`.
// More queries than one pass of the CUDA-core scan takes -> tensor-core scan (when the shape
// fits it): 32 queries per pass over the corpus instead of 8.
inline bool dense_use_tc(const anr_ctx* ctx, const anr_dense* ix, int nq, int k) {
  return nq > 8 && k <= kMaxFusedK && dense_tc_supported(ctx->dp, ix->n, ix->ld, k);
}
// Batches of more than 32 queries: the tiled GEMM (one pass over the corpus for up to 1024 queries).
inline bool dense_shadow_wanted(const anr_dense* ix) {
  static const bool env = getenv("ANR_TC_BF16") != nullptr && atoi(getenv("ANR_TC_BF16")) != 0;
  return (ix->want_shadow || env) && ix->ld % 64 == 0;
}
// Without a shadow copy the CUDA-core fp32 scan (<= 8 queries) and the 32-query tf32 scan read
// the same bytes as the GEMM would, with less set-up; WITH a bf16 shadow the GEMM pass reads half
// the bytes of any fp32 scan, so every batch size takes it (a 64-query tile, zero-padded).
inline int dense_gemm_min_queries(const anr_dense* ix) {
  static const int v = getenv("ANR_GEMM_MIN_QUERIES") ? atoi(getenv("ANR_GEMM_MIN_QUERIES")) : 33;
  static const int vs =
      getenv("ANR_GEMM_MIN_QUERIES_SHADOW") ? atoi(getenv("ANR_GEMM_MIN_QUERIES_SHADOW")) : 1;
  return dense_shadow_wanted(ix) ? vs : v;
}
inline bool dense_use_gemm(const anr_ctx* ctx, const anr_dense* ix, int nq, int k) {
  return nq >= dense_gemm_min_queries(ix) && k <= kMaxFusedK &&
         dense_gemm_supported(ctx->dp, ix->n, ix->ld, k, dense_shadow_wanted(ix));
}
inline int dense_gemm_total_padded(int nq) {
  const int g = dense_gemm_max_queries();
  return nq / g * g + (nq % g ? dense_gemm_padded_queries(nq % g) : 0);
}
inline int dense_padded_queries(const anr_ctx* ctx, const anr_dense* ix, int nq, int k) {
  if (dense_use_gemm(ctx, ix, nq, k)) return dense_gemm_total_padded(nq);
  if (dense_use_tc(ctx, ix, nq, k)) {
    const int p = 2 * dense_tc_queries_per_pass();   // the pair pass reads 64 query rows
    return (nq + p - 1) / p * p;
  }
  return pad_queries(nq, std::max(dense_group(ctx, ix, k), 1));
}

// Shared memory the dense GEMM pass of a hybrid step leaves free on every SM for the kernels of the
// BM25 path beside it: its stage kernels hold <= 4 KB of static shared memory, the rerun pair for
// flagged queries (tiled scan + final top-k) 34 + 18 KB.
constexpr int kBesideBm25Smem = 56 * 1024;

// Does a hybrid step of nq queries take the gated schedule?  Only the single-CTA 64-query GEMM
// pass has a gate (the wider tiles fill the SM on their own).  OPT-IN (ANR_HYBRID_GATE=1).
// Measured (profiles/r2_call10_*, 1M x 1024 + 1M docs, batch 64): the dense main kernel then
// starts at 38 instead of 61 us into the step, but the folded BM25 launch runs 466 us beside it
// (sample launch 60 + main launch 425 us in the default schedule): 0.532 against 0.539 ms per
// step -- and a folded launch classifies head terms against a bound that moves while it runs, so
// the last bit of a BM25 score is no longer repeatable.  Not worth it; the default stays the
// two-launch schedule, whose main launch sees the sample launch's bounds.
inline bool hybrid_gated(const anr_ctx* ctx, const anr_dense* ix, int nq, int k) {
  static const bool on = getenv("ANR_HYBRID_GATE") && atoi(getenv("ANR_HYBRID_GATE")) != 0;
  return on && nq <= 64 && dense_use_gemm(ctx, ix, nq, k);
}

size_t dense_ws_bytes(const anr_ctx* ctx, const anr_dense* ix, int nq, int k) {
  const int gmax = std::max(dense_group(ctx, ix, k), 1);
  const int nqp = pad_queries(nq, gmax);
  if (dense_use_gemm(ctx, ix, nq, k))
    return padded(dense_gemm_scratch_bytes(ctx->dp, ix->n, ix->ld, nq, k)) +
           2 * padded(static_cast<size_t>(nq) * 4) +
           padded(static_cast<size_t>(nq) * dense_scan_max_grid(ctx->dp) * k * 8) + 4096;
  if (dense_use_tc(ctx, ix, nq, k))
    return padded(dense_tc_cand_keys(ctx->dp, ix->n, k) * 8) + 2 * padded(static_cast<size_t>(nq) * 4) +
           padded(static_cast<size_t>(nq) * dense_scan_max_grid(ctx->dp) * k * 8) + 2048;
  if (k <= kMaxFusedK)
    return padded(static_cast<size_t>(nqp) * dense_scan_max_grid(ctx->dp) * k * 8) + 256;
  const int64_t n_pow2 = next_pow2(static_cast<int>(std::max<int64_t>(ix->n, 2)));
  const int group = std::min(nqp, gmax);
  return padded(static_cast<size_t>(group) * n_pow2 * 8) + 256;
}

// ev_pre_main (nullable) is recorded when the pass' first long kernel is next in line on `stream`.
int dense_pipeline(anr_ctx* ctx, const anr_dense* ix, const float* q_dev, int nq, int k,
                   const uint32_t* mask_dev, Arena& arena, const TopkOut& out,
                   cudaStream_t stream, cudaEvent_t ev_pre_main = nullptr,
                   DenseGate* gate = nullptr) {
  const int gmax = dense_group(ctx, ix, k);
  if (gmax < 1) return fail(ANR_ERR_UNSUPPORTED, "embedding rows too long for the scan kernel");
  const bool gemm = dense_use_gemm(ctx, ix, nq, k);
  if (gemm || dense_use_tc(ctx, ix, nq, k)) {
    std::lock_guard<std::mutex> lock(ix->lazy);
    if (!ix->norm_valid) {  // once per index: the error bound needs max |row|
      float* d_norm = arena.take<float>(1);
      ANR_CUDA(launch_row_norm_max(ix->emb, ix->n, ix->ld, d_norm, stream));
      ANR_CUDA(cudaMemcpyAsync(&ix->norm_max, d_norm, 4, cudaMemcpyDeviceToHost, stream));
      ANR_CUDA(cudaStreamSynchronize(stream));
      ix->norm_valid = true;
    }
  }
  if (gemm) {
    if (dense_shadow_wanted(ix)) {  // once per index: the bf16 operand copy
      std::lock_guard<std::mutex> lock(ix->lazy);
      if (!ix->shadow) {
        void* sh = nullptr;
        ANR_CUDA(cudaMalloc(&sh, dense_shadow_bytes(ix->n, ix->ld)));
        cudaError_t e = launch_dense_shadow_fill(ix->emb, sh, ix->n, ix->ld, stream);
        // other streams may use the copy as soon as the pointer is published
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
          cudaFree(sh);
          return fail_cuda("bf16 shadow copy", e);
        }
        ix->shadow = sh;
      }
    }
    unsigned char* scratch =
        arena.take<unsigned char>(dense_gemm_scratch_bytes(ctx->dp, ix->n, ix->ld, nq, k));
    int32_t* flags = arena.take<int32_t>(static_cast<size_t>(nq));
    int32_t* n_flagged = arena.take<int32_t>(1);
    int32_t* flagged = arena.take<int32_t>(static_cast<size_t>(nq));
    const int per = dense_gemm_max_queries();
    // one launch group: its rescoring kernel lists the flagged queries itself (no compaction launch)
    const bool fold_flags = nq <= per && nq <= 1024;   // (one flag per thread of that last CTA)
    for (int q0 = 0; q0 < nq; q0 += per) {
      TopkOut o = out;
      if (o.keys) o.keys += q0 * out.stride_q;
      if (o.scores) o.scores += q0 * out.stride_q;
      if (o.ids) o.ids += q0 * out.stride_q;
      if (o.counts) o.counts += q0 * out.count_stride;
      ProfileScope prof(ctx, 2, stream);
      EventPair* ev = profile_take(ctx, 0);
      EventPair tl_pair;   // timeline: the first group's main kernel
      if (!ev && ctx->timeline && q0 == 0) {
        for (int id : {4, 5})
          if (!ctx->tl[id] && cudaEventCreate(&ctx->tl[id]) != cudaSuccess) cudaGetLastError();
        if (ctx->tl[4] && ctx->tl[5]) {
          tl_pair.start = ctx->tl[4];
          tl_pair.stop = ctx->tl[5];
          ctx->tl_set[4] = ctx->tl_set[5] = true;
          ev = &tl_pair;
        }
      }
      ANR_CUDA(launch_dense_gemm(ctx->dp, ix->emb, ix->shadow, ix->n, ix->ld,
                                 q_dev + static_cast<size_t>(q0) * ix->ld, std::min(per, nq - q0), k,
                                 mask_dev, ix->norm_max, scratch, o, flags + q0,
                                 ev ? ev->start : nullptr, ev ? ev->stop : nullptr, stream,
                                 q0 == 0 ? ev_pre_main : nullptr, q0 == 0 ? gate : nullptr,
                                 fold_flags ? n_flagged : nullptr, fold_flags ? flagged : nullptr));
    }
    tl_mark(ctx, 6, stream);
    const int fb_grid = dense_scan_flagged_grid(ctx->dp, ix->n, ix->ld, k);
    const int64_t fb_stride = static_cast<int64_t>(fb_grid) * k;
    uint64_t* fb_cand = arena.take<uint64_t>(static_cast<size_t>(nq) * fb_stride);
    if (!fold_flags) ANR_CUDA(launch_compact_flags(flags, nq, n_flagged, flagged, stream));
    ctx->last_dense_flagged = n_flagged;
    ANR_CUDA(launch_dense_scan_flagged(ctx->dp, ix->emb, ix->n, ix->ld, q_dev, n_flagged, flagged, k,
                                       mask_dev, fb_cand, fb_stride, stream));
    ANR_CUDA(launch_topk_final_flagged(fb_cand, fb_stride, static_cast<int>(fb_stride), nq, k, out,
                                       n_flagged, flagged, stream));
    return ANR_OK;
  }
  if (ev_pre_main) ANR_CUDA(cudaEventRecord(ev_pre_main, stream));   // (paths without a pre-pass)
  if (dense_use_tc(ctx, ix, nq, k)) {
    const int per = dense_tc_queries_per_pass();
    uint64_t* tc_cand = arena.take<uint64_t>(dense_tc_cand_keys(ctx->dp, ix->n, k));
    int32_t* flags = arena.take<int32_t>(static_cast<size_t>(nq));
    const bool pair = dense_tc_pair_enabled() && (ctx->dp.sm_count % 2) == 0;
    for (int q0 = 0; q0 < nq;) {
      TopkOut o = out;
      if (o.keys) o.keys += q0 * out.stride_q;
      if (o.scores) o.scores += q0 * out.stride_q;
      if (o.ids) o.ids += q0 * out.stride_q;
      if (o.counts) o.counts += q0 * out.count_stride;
      ProfileScope prof(ctx, 2, stream);
      EventPair* ev = profile_take(ctx, 0);
      if (pair && nq - q0 > per) {   // 33..64 queries left: one pass of the CTA-pair kernel
        ANR_CUDA(launch_dense_tc_pair(ctx->dp, ix->emb, ix->n, ix->ld,
                                      q_dev + static_cast<size_t>(q0) * ix->ld,
                                      std::min(2 * per, nq - q0), k, mask_dev, ix->norm_max,
                                      tc_cand, o, flags + q0, ev ? ev->start : nullptr,
                                      ev ? ev->stop : nullptr, stream));
        q0 += 2 * per;
      } else {
        ANR_CUDA(launch_dense_tc(ctx->dp, ix->emb, ix->n, ix->ld,
                                 q_dev + static_cast<size_t>(q0) * ix->ld, std::min(per, nq - q0),
                                 k, mask_dev, ix->norm_max, tc_cand, o, flags + q0,
                                 ev ? ev->start : nullptr, ev ? ev->stop : nullptr, stream));
        q0 += per;
      }
    }
    // Queries whose candidate lists could not prove exactness go through the exact scan.  The
    // list is compacted and consumed on the device (both launches return at once when it is
    // empty), so the call stays asynchronous and capturable in a CUDA graph.
    const int fb_grid = dense_scan_flagged_grid(ctx->dp, ix->n, ix->ld, k);
    const int64_t fb_stride = static_cast<int64_t>(fb_grid) * k;
    int32_t* n_flagged = arena.take<int32_t>(1);
    int32_t* flagged = arena.take<int32_t>(static_cast<size_t>(nq));
    uint64_t* fb_cand = arena.take<uint64_t>(static_cast<size_t>(nq) * fb_stride);
    ANR_CUDA(launch_compact_flags(flags, nq, n_flagged, flagged, stream));
    ctx->last_dense_flagged = n_flagged;
    ANR_CUDA(launch_dense_scan_flagged(ctx->dp, ix->emb, ix->n, ix->ld, q_dev, n_flagged, flagged, k,
                                       mask_dev, fb_cand, fb_stride, stream));
    ANR_CUDA(launch_topk_final_flagged(fb_cand, fb_stride, static_cast<int>(fb_stride), nq, k, out,
                                       n_flagged, flagged, stream));
    return ANR_OK;
  }
  const int nqp = pad_queries(nq, gmax);
  if (k <= kMaxFusedK) {
    const int64_t stride = static_cast<int64_t>(dense_scan_max_grid(ctx->dp)) * k;
    uint64_t* cand = arena.take<uint64_t>(static_cast<size_t>(nqp) * stride);
    int grid = 0;
    for (int q0 = 0; q0 < nqp; q0 += gmax) {
      const int g = std::min(gmax, nqp - q0);
      ProfileScope prof(ctx, 0, stream);
      ANR_CUDA(launch_dense_scan_topk(ctx->dp, ix->emb, ix->n, ix->ld,
                                      q_dev + static_cast<size_t>(q0) * ix->ld, g, k, mask_dev,
                                      cand + q0 * stride, stride, &grid, stream));
    }
    const int m = grid * k;
    ANR_CUDA(launch_topk_final(cand, stride, m, m, 0, nq, k, out, stream));
    return ANR_OK;
  }
  // k > kMaxFusedK: materialise every key, sort, emit (full-ranking path)
  if (ix->n > (1ll << 30)) return fail(ANR_ERR_UNSUPPORTED, "k > 128 needs n <= 2^30");
  const int64_t n_pow2 = next_pow2(static_cast<int>(std::max<int64_t>(ix->n, 2)));
  const int group = std::min(nqp, gmax);
  uint64_t* keys = arena.take<uint64_t>(static_cast<size_t>(group) * n_pow2);
  for (int q0 = 0; q0 < nq; q0 += group) {
    const int real = std::min(group, nq - q0);
    ANR_CUDA(launch_dense_scan_all(ctx->dp, ix->emb, ix->n, ix->ld,
                                   q_dev + static_cast<size_t>(q0) * ix->ld, group, mask_dev, keys,
                                   n_pow2, stream));
    ANR_CUDA(launch_zero_tail(keys, n_pow2, ix->n, n_pow2, real, stream));
    ANR_CUDA(launch_sort_desc(keys, n_pow2, n_pow2, real, stream));
    TopkOut o = out;
    if (o.keys) o.keys += q0 * out.stride_q;
    if (o.scores) o.scores += q0 * out.stride_q;
    if (o.ids) o.ids += q0 * out.stride_q;
    if (o.counts) o.counts += q0 * out.count_stride;
    ANR_CUDA(launch_emit_sorted(keys, n_pow2, n_pow2, real, k, o, stream));
  }
  return ANR_OK;
}

// ---- BM25 top-k pipeline on device buffers ---------------------------------------------
Bm25View bm25_view(const anr_bm25* ix);

// Head terms of the pruned scan: df >= n_docs / div (ANR_BM25_HEAD_DIV, default 4), the
// kBm25MaxHead most frequent at most, within a memory budget of a quarter of the postings' size
// or 1 GB, whichever is larger.
int bm25_ensure_heads(const anr_bm25* ix, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(ix->lazy);
  if (ix->head_built && ix->head_filled) return ANR_OK;
  if (!ix->head_built) {
    std::vector<int64_t> tp(static_cast<size_t>(ix->n_terms) + 1);
    ANR_CUDA(cudaMemcpyAsync(tp.data(), ix->term_ptr, tp.size() * 8, cudaMemcpyDeviceToHost, stream));
    ANR_CUDA(cudaStreamSynchronize(stream));
    // (1M docs, batch 64: df >= N/4 -> 0.254 ms per scan, N/8 -> 0.269, N/16 -> 0.269; profiles/r2_call2_*)
    static const int div = getenv("ANR_BM25_HEAD_DIV") ? std::max(atoi(getenv("ANR_BM25_HEAD_DIV")), 1) : 4;
    const int64_t min_df = std::max<int64_t>(ix->n_docs / div, 1);
    std::vector<std::pair<int64_t, int32_t>> heads;
    for (int32_t t = 0; t < ix->n_terms; ++t) {
      const int64_t df = tp[t + 1] - tp[t];
      if (df >= min_df) heads.emplace_back(-df, t);
    }
    std::sort(heads.begin(), heads.end());
    {  // the pruning bound needs non-negative contributions: an index with a negative idf
       // (possible only when BM25Okapi's average idf is negative) is scanned unpruned
      std::vector<float> idf_h(static_cast<size_t>(std::max(ix->n_terms, 1)), 0.f);
      ANR_CUDA(cudaMemcpyAsync(idf_h.data(), ix->idf, static_cast<size_t>(ix->n_terms) * 4,
                               cudaMemcpyDeviceToHost, stream));
      ANR_CUDA(cudaStreamSynchronize(stream));
      for (int32_t t = 0; t < ix->n_terms; ++t)
        if (idf_h[t] < 0.f) { heads.clear(); break; }
    }
    const int64_t ld = bm25_head_ld(ix->n_docs);
    const size_t budget = std::max<size_t>(static_cast<size_t>(ix->nnz) * 2, static_cast<size_t>(1) << 30);
    size_t n_head = std::min<size_t>(heads.size(), kBm25MaxHead);
    n_head = std::min<size_t>(n_head, budget / (static_cast<size_t>(ld) * 4));
    std::vector<uint8_t> slot(static_cast<size_t>(std::max(ix->n_terms, 1)), 0xff);
    std::vector<int32_t> terms(std::max<size_t>(n_head, 1), 0);
    for (size_t h = 0; h < n_head; ++h) {
      slot[heads[h].second] = static_cast<uint8_t>(h);
      terms[h] = heads[h].second;
    }
    ANR_CUDA(cudaMalloc(&ix->head_slot, slot.size()));
    ANR_CUDA(cudaMalloc(&ix->head_terms, terms.size() * 4));
    ANR_CUDA(cudaMalloc(&ix->head_w, std::max<size_t>(n_head, 1) * ld * 4));
    ANR_CUDA(cudaMalloc(&ix->head_max, std::max<size_t>(n_head, 1) * 4));
    ANR_CUDA(cudaMemcpyAsync(ix->head_slot, slot.data(), slot.size(), cudaMemcpyHostToDevice, stream));
    ANR_CUDA(cudaMemcpyAsync(ix->head_terms, terms.data(), terms.size() * 4, cudaMemcpyHostToDevice,
                             stream));
    ANR_CUDA(cudaStreamSynchronize(stream));   // the host vectors go out of scope
    ix->head_ld = ld;
    ix->n_head = static_cast<int32_t>(n_head);
    ix->head_built = true;
  }
  ANR_CUDA(launch_bm25_head_fill(bm25_view(ix), ix->head_terms, ix->n_head, ix->head_w,
                                 ix->head_max, ix->head_ld, stream));
  ANR_CUDA(cudaStreamSynchronize(stream));   // other streams may read the rows once the flag is set
  ix->head_filled = true;
  return ANR_OK;
}

// Candidate-driven top-k (anr_bm25_ms.cu) is the default for k <= 128; ANR_BM25_MAXSCORE=0 keeps
// the tiled scan.  Tiny indices and batches beyond one grid dimension stay on the tiled scan.
// (The two knobs are read on every call: the GPU tests switch paths inside one process.)
bool bm25_ms_wanted(const anr_bm25* ix, int nq, int k) {
  const char* e = getenv("ANR_BM25_MAXSCORE");
  const bool on = !(e && atoi(e) == 0) && getenv("ANR_DISABLE_BM25_PRUNE") == nullptr;
  return on && k <= kMaxFusedK && ix->n_docs >= 1024 && ix->nnz > 0 && nq <= 65535;
}
// Once per index (and after a reweighting): per-term largest weight + "is any idf negative".
int bm25_ensure_ms(const anr_bm25* ix, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(ix->lazy);
  if (ix->ms_ready) return ANR_OK;
  const size_t nt = static_cast<size_t>(std::max(ix->n_terms, 1));
  if (!ix->term_maxw) ANR_CUDA(cudaMalloc(&ix->term_maxw, nt * 4 + 16));
  int32_t* neg_dev = reinterpret_cast<int32_t*>(ix->term_maxw + nt);
  ANR_CUDA(launch_bm25_term_max(bm25_view(ix), ix->term_maxw, neg_dev, stream));
  int32_t neg = 0;
  ANR_CUDA(cudaMemcpyAsync(&neg, neg_dev, 4, cudaMemcpyDeviceToHost, stream));
  if (!ix->bkt_off) {   // the tables follow the postings' positions only: built once per index
    std::vector<int64_t> tp(static_cast<size_t>(ix->n_terms) + 1);
    ANR_CUDA(cudaMemcpyAsync(tp.data(), ix->term_ptr, tp.size() * 8, cudaMemcpyDeviceToHost, stream));
    ANR_CUDA(cudaStreamSynchronize(stream));
    std::vector<int64_t> off(nt, 0);
    std::vector<uint8_t> shift(nt, 0xff);
    int64_t total = 0;
    for (int32_t t = 0; t < ix->n_terms; ++t) {
      const int sh = bm25_bucket_shift(tp[t + 1] - tp[t], ix->n_docs);
      shift[t] = static_cast<uint8_t>(sh);
      off[t] = total;
      total += bm25_bucket_entries(sh, ix->n_docs);
    }
    ANR_CUDA(cudaMalloc(&ix->bkt_off, nt * 8));
    ANR_CUDA(cudaMalloc(&ix->bkt_shift, nt));
    ANR_CUDA(cudaMalloc(&ix->bkt, static_cast<size_t>(std::max<int64_t>(total, 1)) * 4));
    ANR_CUDA(cudaMemcpyAsync(ix->bkt_off, off.data(), nt * 8, cudaMemcpyHostToDevice, stream));
    ANR_CUDA(cudaMemcpyAsync(ix->bkt_shift, shift.data(), nt, cudaMemcpyHostToDevice, stream));
    ANR_CUDA(launch_bm25_bucket_fill(bm25_view(ix), ix->bkt_off, ix->bkt_shift, ix->bkt, stream));
  }
  ANR_CUDA(cudaStreamSynchronize(stream));   // host vectors go out of scope; other streams may read
  ix->neg_idf = neg != 0;
  ix->ms_ready = true;
  return ANR_OK;
}

size_t bm25_ws_bytes(const anr_ctx* ctx, const anr_bm25* ix, int nq, int k) {
  if (k <= kMaxFusedK) {
    const Bm25Plan plan = bm25_make_plan(ctx->dp, ix->n_docs, nq, k, false);
    const size_t slots = static_cast<size_t>(plan.n_tiles);
    size_t ms = 0;
    if (bm25_ms_wanted(ix, nq, k))
      ms = padded(bm25_ms_scratch_bytes(nq)) + padded(static_cast<size_t>(nq) * kMsSurvivors * 8) +
           padded(static_cast<size_t>(nq) * 4) + 1024;
    return padded(static_cast<size_t>(nq) * slots * k * 8) +
           padded(static_cast<size_t>(nq) * 4) + 512 + ms;
  }
  const int64_t n_pow2 = next_pow2(std::max(ix->n_docs, 2));
  return padded(static_cast<size_t>(std::min(nq, 8)) * n_pow2 * 8) + 256;
}

Bm25View bm25_view(const anr_bm25* ix) {
  Bm25View v;
  v.term_ptr = ix->term_ptr;
  v.post_doc = ix->post_doc;
  v.post_w = ix->post_w;
  v.idf = ix->idf;
  v.n_terms = ix->n_terms;
  v.n_docs = ix->n_docs;
  v.nnz = ix->nnz;
  return v;
}

// State of a BM25 top-k scan issued in two phases around the dense pass of a hybrid query.
struct Bm25Run {
  bool active = false;
  bool ms_done = false;   // the candidate-driven path finished the whole search in phase 1
  Bm25Plan plan;
  Bm25HeadView hd;
  uint64_t* cand = nullptr;
  float* theta = nullptr;
  int64_t stride = 0;
};

// phase 0: the whole search.  Hybrid queries split it: phase 1 = set-up + the short sample launch
// of the pruned scan (state kept in *run), phase 2 = the main launch + final top-k, enqueued once
// the dense pass' main kernel is next in line on its own stream (see anr_hybrid_search).
int bm25_pipeline(anr_ctx* ctx, const anr_bm25* ix, const int32_t* terms_dev,
                  const int32_t* offsets_dev, int nq, int k, const uint32_t* mask_dev,
                  Arena& arena, const TopkOut& out, cudaStream_t stream,
                  bool beside_dense = false, int phase = 0, Bm25Run* run = nullptr) {
  const Bm25View v = bm25_view(ix);
  if (phase == 2 && run && run->ms_done) return ANR_OK;
  if (bm25_ms_wanted(ix, nq, k)) {
    if (int rc = bm25_ensure_ms(ix, stream)) return rc;
    if (!ix->neg_idf) {
      // ---- candidate-driven top-k: plan | stage 1 | theta + required lists | stage 2 | final top-k,
      //      then the exhaustive tiled scan for the (usually zero) flagged queries ----
      Bm25HeadView hd;
      if (ix->n_docs >= 8192) {
        if (int rc = bm25_ensure_heads(ix, stream)) return rc;
        hd.slot_of = ix->head_slot;
        hd.head_w = ix->head_w;
        hd.head_max = ix->head_max;
        hd.head_ld = ix->head_ld;
        hd.n_head = ix->n_head;
      }
      unsigned char* scratch = arena.take<unsigned char>(bm25_ms_scratch_bytes(nq));
      uint64_t* surv = arena.take<uint64_t>(static_cast<size_t>(nq) * kMsSurvivors);
      int32_t* flagged = arena.take<int32_t>(static_cast<size_t>(nq));
      int32_t* n_flagged = arena.take<int32_t>(1);
      const Bm25Plan plan = bm25_make_plan(ctx->dp, ix->n_docs, nq, k, false);
      if (plan.smem_bytes > ctx->dp.max_smem_optin)
        return fail(ANR_ERR_UNSUPPORTED, "bm25 tile does not fit in shared memory");
      const int64_t fb_stride = static_cast<int64_t>(plan.n_tiles) * k;
      uint64_t* fb_cand = arena.take<uint64_t>(static_cast<size_t>(nq) * fb_stride);
      TopkOut live = out;
      live.q_offsets = offsets_dev;   // a query without terms gets no result (search_engine.py:216-217)
      tl_mark(ctx, 1, stream);
      {
        ProfileScope prof(ctx, 1, stream);
        MsIndexView mx;
        mx.term_maxw = ix->term_maxw;
        mx.bkt_off = ix->bkt_off;
        mx.bkt_shift = ix->bkt_shift;
        mx.bkt = ix->bkt;
        cudaEvent_t marks[4] = {nullptr, nullptr, nullptr, nullptr};
        if (ctx->timeline)
          for (int i = 0; i < 4; ++i) {
            const int id = 12 + i;
            if (!ctx->tl[id] && cudaEventCreate(&ctx->tl[id]) != cudaSuccess) cudaGetLastError();
            marks[i] = ctx->tl[id];
            ctx->tl_set[id] = marks[i] != nullptr;
          }
        ANR_CUDA(launch_bm25_maxscore(ctx->dp, v, hd, mx, terms_dev, offsets_dev, nq, k, mask_dev,
                                      scratch, surv, live, n_flagged, flagged, stream,
                                      beside_dense || ctx->beside_dense, marks));
      }
      tl_mark(ctx, 2, stream);
      tl_mark(ctx, 8, stream);
      tl_mark(ctx, 9, stream);
      if (getenv("ANR_MS_DEBUG")) {   // per-query statistics of the path (profiling runs only: synchronises)
        struct Q { int64_t s2_total; int32_t n; float theta; int32_t s1_total, n_surv, flag; float tot; };
        std::vector<Q> h(static_cast<size_t>(nq));
        ANR_CUDA(cudaStreamSynchronize(stream));
        ANR_CUDA(cudaMemcpy(h.data(), scratch, h.size() * sizeof(Q), cudaMemcpyDeviceToHost));
        int64_t items = 0, surv_n = 0;
        int flagged_n = 0;
        for (const Q& x : h) { items += x.s2_total; surv_n += x.n_surv; flagged_n += x.flag != 0 || x.n_surv < k; }
        fprintf(stderr, "[anr ms] nq %d k %d stage-2 postings %lld survivors %lld flagged %d\n", nq, k,
                static_cast<long long>(items), static_cast<long long>(surv_n), flagged_n);
        for (int i = 0; i < nq && i < 64; ++i)
          fprintf(stderr, "[anr ms]   q%d terms %d theta %.3f tot %.2f s1 %d s2 %lld surv %d flag %d\n", i, h[i].n,
                  h[i].theta, h[i].tot, h[i].s1_total, static_cast<long long>(h[i].s2_total), h[i].n_surv,
                  h[i].flag);
      }
      ctx->last_bm25_flagged = n_flagged;
      ANR_CUDA(launch_bm25_score_listed(v, terms_dev, offsets_dev, nq, k, mask_dev, plan, fb_cand,
                                        fb_stride, flagged, n_flagged, stream));
      ANR_CUDA(launch_topk_final_flagged(fb_cand, fb_stride, static_cast<int>(fb_stride), nq, k, live,
                                         n_flagged, flagged, stream));
      tl_mark(ctx, 10, stream);
      if (run) run->ms_done = true;
      return ANR_OK;
    }
  }
  if (k <= kMaxFusedK) {
    Bm25Run local;
    Bm25Run& r = run ? *run : local;
    if (!(phase == 2 && r.active)) {
      r.plan = bm25_make_plan(ctx->dp, ix->n_docs, nq, k, false);
      r.plan.beside_dense = beside_dense || ctx->beside_dense;
      if (r.plan.smem_bytes > ctx->dp.max_smem_optin)
        return fail(ANR_ERR_UNSUPPORTED, "bm25 tile does not fit in shared memory");
      // safe dynamic pruning over the dense rows of the head terms (corpora of 8192+ documents)
      const bool no_prune = getenv("ANR_DISABLE_BM25_PRUNE") != nullptr;
      r.hd = Bm25HeadView();
      if (!no_prune && ix->n_docs >= 8192 && nq >= 16) {   // (measured: no gain below ~16 queries)
        if (int rc = bm25_ensure_heads(ix, stream)) return rc;
        r.hd.slot_of = ix->head_slot;
        r.hd.head_w = ix->head_w;
        r.hd.head_max = ix->head_max;
        r.hd.head_ld = ix->head_ld;
        r.hd.n_head = ix->n_head;
      }
      r.stride = static_cast<int64_t>(r.plan.n_tiles) * k;
      r.cand = arena.take<uint64_t>(static_cast<size_t>(nq) * r.stride);
      r.theta = arena.take<float>(static_cast<size_t>(nq));
      r.active = true;
    }
    Bm25Plan plan = r.plan;
    plan.phase = phase;
    const Bm25HeadView& hd = r.hd;
    uint64_t* cand = r.cand;
    float* theta = r.theta;
    const int64_t stride = r.stride;
    if (phase == 1) {
      tl_mark(ctx, 1, stream);
      ANR_CUDA(launch_bm25_score_topk(v, hd.n_head > 0 ? &hd : nullptr, terms_dev, offsets_dev, nq,
                                      k, mask_dev, plan, cand, stride, theta, stream));
      tl_mark(ctx, 2, stream);
      return ANR_OK;
    }
    tl_mark(ctx, 8, stream);
    {
      ProfileScope prof(ctx, 1, stream);
      ANR_CUDA(launch_bm25_score_topk(v, hd.n_head > 0 ? &hd : nullptr, terms_dev, offsets_dev, nq,
                                      k, mask_dev, plan, cand, stride, theta, stream));
    }
    tl_mark(ctx, 9, stream);
    const int m = static_cast<int>(stride);
    TopkOut live = out;
    live.q_offsets = offsets_dev;   // a query without terms gets no result (search_engine.py:216-217)
    ANR_CUDA(launch_topk_final(cand, stride, m, m, 0, nq, k, live, stream));
    tl_mark(ctx, 10, stream);
    return ANR_OK;
  }
  if (phase == 1) return ANR_OK;   // the full-ranking path has no sample launch
  if (ix->n_docs > (1 << 30)) return fail(ANR_ERR_UNSUPPORTED, "k > 128 needs n_docs <= 2^30");
  const int64_t n_pow2 = next_pow2(std::max(ix->n_docs, 2));
  const int group = std::min(nq, 8);
  uint64_t* keys = arena.take<uint64_t>(static_cast<size_t>(group) * n_pow2);
  for (int q0 = 0; q0 < nq; q0 += group) {
    const int real = std::min(group, nq - q0);
    const Bm25Plan plan = bm25_make_plan(ctx->dp, ix->n_docs, real, 1, true);
    ANR_CUDA(launch_bm25_score_all(v, terms_dev, offsets_dev + q0, real, mask_dev, plan, keys,
                                   n_pow2, stream));
    ANR_CUDA(launch_zero_tail(keys, n_pow2, ix->n_docs, n_pow2, real, stream));
    ANR_CUDA(launch_sort_desc(keys, n_pow2, n_pow2, real, stream));
    TopkOut o = out;
    if (o.keys) o.keys += q0 * out.stride_q;
    if (o.scores) o.scores += q0 * out.stride_q;
    if (o.ids) o.ids += q0 * out.stride_q;
    if (o.counts) o.counts += q0 * out.count_stride;
    o.q_offsets = offsets_dev + q0;   // a query without terms gets no result
    ANR_CUDA(launch_emit_sorted(keys, n_pow2, n_pow2, real, k, o, stream));
  }
  return ANR_OK;
}

// Stage the query matrix: returns a device [pad_queries(nq), ld] zero-padded copy.
int stage_queries(const anr_dense* ix, const float* queries, int nq, int nqp, Arena& arena,
                  cudaStream_t stream, const float** q_dev) {
  // device-resident queries that need no padding are read in place
  if (nqp == nq && ix->ld == ix->d && reinterpret_cast<uintptr_t>(queries) % 16 == 0 &&
      is_device_ptr(queries)) {
    *q_dev = queries;
    return ANR_OK;
  }
  float* q = arena.take<float>(static_cast<size_t>(nqp) * ix->ld);
  if (nqp != nq || ix->ld != ix->d)
    ANR_CUDA(cudaMemsetAsync(q, 0, static_cast<size_t>(nqp) * ix->ld * 4, stream));
  if (ix->ld == ix->d) {
    ANR_CUDA(cudaMemcpyAsync(q, queries, static_cast<size_t>(nq) * ix->d * 4, cudaMemcpyDefault,
                             stream));
  } else {
    ANR_CUDA(cudaMemcpy2DAsync(q, static_cast<size_t>(ix->ld) * 4, queries,
                               static_cast<size_t>(ix->d) * 4, static_cast<size_t>(ix->d) * 4, nq,
                               cudaMemcpyDefault, stream));
  }
  *q_dev = q;
  return ANR_OK;
}
size_t stage_queries_bytes(const anr_dense* ix, int nq) {
  return padded(static_cast<size_t>(nq + 256) * ix->ld * 4) + 256;
}

// Stage a bit mask ([ceil(n/32)] words) if it lives on the host.
size_t stage_mask_bytes(const uint32_t* mask, int64_t n) {
  return (mask && !is_device_ptr(mask)) ? padded(static_cast<size_t>((n + 31) / 32) * 4) + 256 : 0;
}
int stage_mask(const uint32_t* mask, int64_t n, Arena& arena, cudaStream_t stream,
               const uint32_t** mask_dev) {
  *mask_dev = mask;
  if (!mask || is_device_ptr(mask)) return ANR_OK;
  const size_t words = static_cast<size_t>((n + 31) / 32);
  uint32_t* m = arena.take<uint32_t>(words);
  ANR_CUDA(cudaMemcpyAsync(m, mask, words * 4, cudaMemcpyHostToDevice, stream));
  *mask_dev = m;
  return ANR_OK;
}

// Stage the CSR query terms.  Host arrays are copied; device arrays are used in place.
struct QueryTerms {
  const int32_t* terms = nullptr;
  const int32_t* offsets = nullptr;
};
size_t stage_terms_bytes(const int32_t* q_terms, const int32_t* q_offsets, int nq) {
  size_t need = 0;
  if (!is_device_ptr(q_offsets)) {
    need += padded(static_cast<size_t>(nq + 1) * 4) + 256;
    if (!is_device_ptr(q_terms)) need += padded(static_cast<size_t>(std::max(q_offsets[nq], 1)) * 4) + 256;
  }
  return need;
}
int stage_terms(const int32_t* q_terms, const int32_t* q_offsets, int nq, Arena& arena,
                cudaStream_t stream, QueryTerms* out) {
  const bool off_dev = is_device_ptr(q_offsets), terms_dev = is_device_ptr(q_terms);
  if (off_dev != terms_dev && q_terms)
    return fail(ANR_ERR_INVALID, "q_terms and q_offsets must both be host or both be device");
  out->terms = q_terms;
  out->offsets = q_offsets;
  if (off_dev) return ANR_OK;
  if (q_offsets[0] != 0) return fail(ANR_ERR_INVALID, "q_offsets[0] must be 0");
  for (int i = 0; i < nq; ++i)
    if (q_offsets[i + 1] < q_offsets[i]) return fail(ANR_ERR_INVALID, "q_offsets must not decrease");
  int32_t* off = arena.take<int32_t>(static_cast<size_t>(nq) + 1);
  ANR_CUDA(cudaMemcpyAsync(off, q_offsets, (static_cast<size_t>(nq) + 1) * 4,
                           cudaMemcpyHostToDevice, stream));
  const int total = q_offsets[nq];
  int32_t* t = arena.take<int32_t>(static_cast<size_t>(std::max(total, 1)));
  if (total > 0)
    ANR_CUDA(cudaMemcpyAsync(t, q_terms, static_cast<size_t>(total) * 4, cudaMemcpyHostToDevice,
                             stream));
  out->terms = t;
  out->offsets = off;
  return ANR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------
// ABI
// ---------------------------------------------------------------------------------
extern "C" {

int anr_abi_version(void) { return ANR_ABI_VERSION; }
const char* anr_last_error(void) { return g_last_error.c_str(); }

int anr_ctx_create(int device, anr_ctx** out) {
  if (!out) return fail(ANR_ERR_INVALID, "anr_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
    cudaGetLastError();
    return fail(ANR_ERR_NO_DEVICE, "no CUDA device: this library has no CPU implementation");
  }
  if (device < 0 || device >= count) return fail(ANR_ERR_INVALID, "device ordinal out of range");
  cudaDeviceProp prop;
  ANR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    char msg[128];
    snprintf(msg, sizeof msg, "device %d is sm_%d%d; the kernels are built for sm_100a only", device,
             prop.major, prop.minor);
    return fail(ANR_ERR_NO_DEVICE, msg);
  }
  anr_ctx* ctx = new (std::nothrow) anr_ctx;
  if (!ctx) return fail(ANR_ERR_OOM, "host allocation failed");
  ctx->dp.device = device;
  ctx->dp.sm_count = prop.multiProcessorCount;
  ctx->dp.max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  DeviceGuard guard(device);
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    // BM25 of a hybrid query runs here under the dense pass: lowest priority, so that pending CTAs
    // of the (bandwidth-bound, persistent) dense kernels are placed first
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    e = cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, lo);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_mid, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_last, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    anr_ctx_destroy(ctx);
    return fail_cuda("stream/event creation", e);
  }
  *out = ctx;
  return ANR_OK;
}

int anr_set_option(const char* key, int32_t value) {
  if (!key) return fail(ANR_ERR_INVALID, "anr_set_option: key is NULL");
  if (strcmp(key, "pdl") == 0) {
    pdl_set_mask(value);
    return ANR_OK;
  }
  return fail(ANR_ERR_INVALID, "anr_set_option: unknown key");
}

int anr_ctx_profile_enable(anr_ctx* ctx, int32_t on) {
  if (!ctx) return fail(ANR_ERR_INVALID, "ctx is NULL");
  ctx->profiling = on != 0;
  return ANR_OK;
}

int anr_ctx_profile_read(anr_ctx* ctx, int32_t kind, double* total_ms, int64_t* launches) {
  if (!ctx || kind < 0 || kind > 2) return fail(ANR_ERR_INVALID, "anr_ctx_profile_read: bad argument");
  DeviceGuard guard(ctx->dp.device);
  ANR_CUDA(cudaDeviceSynchronize());
  double sum = 0.0;
  for (size_t i = 0; i < ctx->used[kind]; ++i) {
    float ms = 0.f;
    ANR_CUDA(cudaEventElapsedTime(&ms, ctx->pool[kind][i].start, ctx->pool[kind][i].stop));
    sum += ms;
  }
  if (total_ms) *total_ms = sum;
  if (launches) *launches = static_cast<int64_t>(ctx->used[kind]);
  ctx->used[kind] = 0;
  return ANR_OK;
}

int anr_ctx_timeline_enable(anr_ctx* ctx, int32_t on) {
  if (!ctx) return fail(ANR_ERR_INVALID, "ctx is NULL");
  ctx->timeline = on != 0;
  for (bool& b : ctx->tl_set) b = false;
  return ANR_OK;
}

int anr_ctx_timeline_read(anr_ctx* ctx, double* offsets_ms) {
  if (!ctx || !offsets_ms) return fail(ANR_ERR_INVALID, "anr_ctx_timeline_read: NULL argument");
  DeviceGuard guard(ctx->dp.device);
  ANR_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < ANR_TIMELINE_MARKS; ++i) {
    offsets_ms[i] = -1.0;
    if (!ctx->tl_set[0] || !ctx->tl_set[i]) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->tl[0], ctx->tl[i]) == cudaSuccess) offsets_ms[i] = ms;
    else cudaGetLastError();
  }
  return ANR_OK;
}

int anr_ctx_destroy(anr_ctx* ctx) {
  if (!ctx) return ANR_OK;
  DeviceGuard guard(ctx->dp.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->side) cudaStreamSynchronize(ctx->side);
  for (auto& pool : ctx->pool)
    for (auto& p : pool) {
      cudaEventDestroy(p.start);
      cudaEventDestroy(p.stop);
    }
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_mid) cudaEventDestroy(ctx->ev_mid);
  if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
  for (cudaEvent_t e : ctx->tl)
    if (e) cudaEventDestroy(e);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return ANR_OK;
}

int anr_ctx_set_beside_dense(anr_ctx* ctx, int32_t enable) {
  if (!ctx) return fail(ANR_ERR_INVALID, "ctx is NULL");
  ctx->beside_dense = enable != 0;
  return ANR_OK;
}

int anr_ctx_sync(anr_ctx* ctx) {
  if (!ctx) return fail(ANR_ERR_INVALID, "ctx is NULL");
  DeviceGuard guard(ctx->dp.device);
  ANR_CUDA(cudaStreamSynchronize(ctx->stream));
  return ANR_OK;
}

int anr_ctx_last_rerun(anr_ctx* ctx, int32_t* dense_queries, int32_t* bm25_queries) {
  if (!ctx) return fail(ANR_ERR_INVALID, "anr_ctx_last_rerun: ctx is NULL");
  DeviceGuard guard(ctx->dp.device);
  ANR_CUDA(cudaDeviceSynchronize());
  int32_t d = -1, b = -1;
  if (ctx->last_dense_flagged)
    ANR_CUDA(cudaMemcpy(&d, ctx->last_dense_flagged, 4, cudaMemcpyDeviceToHost));
  if (ctx->last_bm25_flagged)
    ANR_CUDA(cudaMemcpy(&b, ctx->last_bm25_flagged, 4, cudaMemcpyDeviceToHost));
  if (dense_queries) *dense_queries = d;
  if (bm25_queries) *bm25_queries = b;
  return ANR_OK;
}

int anr_ctx_info(anr_ctx* ctx, int32_t* sm_count, int64_t* hbm_total, int64_t* hbm_free) {
  if (!ctx) return fail(ANR_ERR_INVALID, "ctx is NULL");
  DeviceGuard guard(ctx->dp.device);
  size_t f = 0, t = 0;
  ANR_CUDA(cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = ctx->dp.sm_count;
  if (hbm_total) *hbm_total = static_cast<int64_t>(t);
  if (hbm_free) *hbm_free = static_cast<int64_t>(f);
  return ANR_OK;
}

// ---- dense index --------------------------------------------------------------
int anr_dense_create(anr_ctx* ctx, const float* emb, int64_t n, int32_t d, int32_t borrow,
                     anr_dense** out) {
  if (!ctx || !out) return fail(ANR_ERR_INVALID, "anr_dense_create: NULL argument");
  *out = nullptr;
  if (n < 0 || d < 1) return fail(ANR_ERR_INVALID, "anr_dense_create: bad shape");
  if (n >= (1ll << 32) - 1) return fail(ANR_ERR_UNSUPPORTED, "more than 2^32-2 rows per index");
  DeviceGuard guard(ctx->dp.device);
  anr_dense* ix = new (std::nothrow) anr_dense;
  if (!ix) return fail(ANR_ERR_OOM, "host allocation failed");
  ix->device = ctx->dp.device;
  ix->n = n;
  ix->d = d;
  ix->ld = (d + 3) / 4 * 4;
  if (borrow) {
    if (!emb || !is_device_ptr(emb) || ix->ld != d ||
        (reinterpret_cast<uintptr_t>(emb) & 15) != 0) {
      delete ix;
      return fail(ANR_ERR_INVALID,
                  "borrow needs a 16-byte aligned device pointer and d % 4 == 0");
    }
    ix->emb = const_cast<float*>(emb);
    ix->owned = false;
    *out = ix;
    return ANR_OK;
  }
  const size_t bytes = std::max<size_t>(static_cast<size_t>(n) * ix->ld * 4, 16);
  cudaError_t e = cudaMalloc(&ix->emb, bytes);
  if (e != cudaSuccess) {
    delete ix;
    return fail_cuda("cudaMalloc(embedding matrix)", e);
  }
  if (emb && n > 0) {
    e = copy_rows(ix->emb, ix->ld, emb, d, n, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cudaFree(ix->emb);
      delete ix;
      return fail_cuda("upload of the embedding matrix", e);
    }
  }
  *out = ix;
  return ANR_OK;
}

int anr_dense_upload(anr_ctx* ctx, anr_dense* index, int64_t row0, const float* rows,
                     int64_t n_rows) {
  if (!ctx || !index || !rows) return fail(ANR_ERR_INVALID, "anr_dense_upload: NULL argument");
  if (!index->owned) return fail(ANR_ERR_INVALID, "anr_dense_upload: borrowed index");
  if (row0 < 0 || n_rows < 0 || row0 + n_rows > index->n)
    return fail(ANR_ERR_INVALID, "anr_dense_upload: row range out of bounds");
  DeviceGuard guard(ctx->dp.device);
  std::lock_guard<std::mutex> lock(index->lazy);
  index->norm_valid = false;
  if (index->shadow) {  // rebuilt on the next tensor-core search
    ANR_CUDA(cudaDeviceSynchronize());
    cudaFree(index->shadow);
    index->shadow = nullptr;
  }
  ANR_CUDA(copy_rows(index->emb + static_cast<size_t>(row0) * index->ld, index->ld, rows, index->d,
                     n_rows, ctx->stream));
  // the source may be a pageable/pinned buffer the caller is about to reuse
  ANR_CUDA(cudaStreamSynchronize(ctx->stream));
  return ANR_OK;
}

int anr_dense_destroy(anr_dense* index) {
  if (!index) return ANR_OK;
  DeviceGuard guard(index->device);
  if (index->owned && index->emb) cudaFree(index->emb);
  if (index->shadow) cudaFree(index->shadow);
  delete index;
  return ANR_OK;
}

int anr_dense_set_shadow(anr_dense* index, int32_t enable) {
  if (!index) return fail(ANR_ERR_INVALID, "index is NULL");
  DeviceGuard guard(index->device);
  std::lock_guard<std::mutex> lock(index->lazy);
  index->want_shadow = enable != 0;
  if (!enable && index->shadow) {
    ANR_CUDA(cudaDeviceSynchronize());
    cudaFree(index->shadow);
    index->shadow = nullptr;
  }
  return ANR_OK;
}

int anr_dense_invalidate(anr_dense* index) {
  if (!index) return fail(ANR_ERR_INVALID, "index is NULL");
  DeviceGuard guard(index->device);
  std::lock_guard<std::mutex> lock(index->lazy);
  index->norm_valid = false;
  if (index->shadow) {   // rebuilt from the current rows by the next tensor-core search
    ANR_CUDA(cudaDeviceSynchronize());
    cudaFree(index->shadow);
    index->shadow = nullptr;
  }
  return ANR_OK;
}

int anr_dense_shape(const anr_dense* index, int64_t* n, int32_t* d) {
  if (!index) return fail(ANR_ERR_INVALID, "index is NULL");
  if (n) *n = index->n;
  if (d) *d = index->d;
  return ANR_OK;
}

static int dense_search_impl(anr_ctx* ctx, const anr_dense* index, const float* queries,
                             int32_t nq, int32_t k, const uint32_t* row_mask, int64_t id_base,
                             uint64_t* out_keys, float* out_scores, int32_t* out_rows,
                             int32_t* out_counts, void* stream_v) {
  if (!ctx || !index || !queries) return fail(ANR_ERR_INVALID, "dense search: NULL argument");
  if (nq < 1 || k < 1) return fail(ANR_ERR_INVALID, "dense search: n_queries and k must be >= 1");
  if (index->device != ctx->dp.device) return fail(ANR_ERR_INVALID, "index lives on another device");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t cells = static_cast<size_t>(nq) * k;
  const size_t need = stage_queries_bytes(index, nq) + stage_mask_bytes(row_mask, index->n) +
                      dense_ws_bytes(ctx, index, nq, k) + out_need(out_keys, cells) +
                      out_need(out_scores, cells) + out_need(out_rows, cells) +
                      out_need(out_counts, nq);
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  OutBuf<uint64_t> o_keys = out_make(arena, out_keys, cells);
  OutBuf<float> o_scores = out_make(arena, out_scores, cells);
  OutBuf<int32_t> o_rows = out_make(arena, out_rows, cells);
  OutBuf<int32_t> o_counts = out_make(arena, out_counts, nq);
  TopkOut out;
  out.keys = o_keys.dev;
  out.scores = o_scores.dev;
  out.ids = o_rows.dev;
  out.counts = o_counts.dev;
  out.stride_q = k;
  out.id_base = id_base;
  if (index->n == 0) {
    if (out.keys) ANR_CUDA(cudaMemsetAsync(out.keys, 0, cells * 8, stream));
    if (out.scores) ANR_CUDA(cudaMemsetAsync(out.scores, 0, cells * 4, stream));
    if (out.ids) ANR_CUDA(cudaMemsetAsync(out.ids, 0xff, cells * 4, stream));
    if (out.counts) ANR_CUDA(cudaMemsetAsync(out.counts, 0, static_cast<size_t>(nq) * 4, stream));
  } else {
    const float* q_dev = nullptr;
    const uint32_t* mask_dev = nullptr;
    if (int rc = stage_queries(index, queries, nq, dense_padded_queries(ctx, index, nq, k), arena, stream,
                              &q_dev)) return rc;
    if (int rc = stage_mask(row_mask, index->n, arena, stream, &mask_dev)) return rc;
    if (int rc = dense_pipeline(ctx, index, q_dev, nq, k, mask_dev, arena, out, stream)) return rc;
  }
  bool any_host = false;
  ANR_CUDA(out_flush(o_keys, stream, &any_host));
  ANR_CUDA(out_flush(o_scores, stream, &any_host));
  ANR_CUDA(out_flush(o_rows, stream, &any_host));
  ANR_CUDA(out_flush(o_counts, stream, &any_host));
  if (any_host) ANR_CUDA(cudaStreamSynchronize(stream));
  return ANR_OK;
}

int anr_dense_search(anr_ctx* ctx, const anr_dense* index, const float* queries,
                     int32_t n_queries, int32_t k, const uint32_t* row_mask, int64_t id_base,
                     float* out_scores, int32_t* out_rows, int32_t* out_counts, void* stream) {
  return dense_search_impl(ctx, index, queries, n_queries, k, row_mask, id_base, nullptr,
                           out_scores, out_rows, out_counts, stream);
}

int anr_dense_search_keys(anr_ctx* ctx, const anr_dense* index, const float* queries,
                          int32_t n_queries, int32_t k, const uint32_t* row_mask, int64_t id_base,
                          uint64_t* out_keys, void* stream) {
  if (!out_keys) return fail(ANR_ERR_INVALID, "out_keys is NULL");
  return dense_search_impl(ctx, index, queries, n_queries, k, row_mask, id_base, out_keys, nullptr,
                           nullptr, nullptr, stream);
}

// ---- BM25 index ---------------------------------------------------------------------
int anr_bm25_create(anr_ctx* ctx, const int64_t* term_ptr, const int32_t* post_doc,
                    const int32_t* post_tf, const int32_t* doc_len, const double* idf,
                    int32_t n_terms, int32_t n_docs, double k1, double b, double avgdl,
                    anr_bm25** out) {
  if (!ctx || !out) return fail(ANR_ERR_INVALID, "anr_bm25_create: NULL argument");
  *out = nullptr;
  if (n_terms < 0 || n_docs < 0) return fail(ANR_ERR_INVALID, "anr_bm25_create: bad shape");
  if (!term_ptr || !idf || (n_docs > 0 && !doc_len))
    return fail(ANR_ERR_INVALID, "anr_bm25_create: NULL array");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = ctx->stream;
  anr_bm25* ix = new (std::nothrow) anr_bm25;
  if (!ix) return fail(ANR_ERR_OOM, "host allocation failed");
  ix->device = ctx->dp.device;
  ix->n_terms = n_terms;
  ix->n_docs = n_docs;
  int32_t* tf_dev = nullptr;
  int32_t* dl_dev = nullptr;
  double* idf64 = nullptr;
  cudaError_t e = cudaSuccess;
  auto cleanup = [&]() {
    if (tf_dev) cudaFree(tf_dev);
    if (dl_dev) cudaFree(dl_dev);
    if (idf64) cudaFree(idf64);
  };
  auto bail = [&](const char* what) {
    cleanup();
    anr_bm25_destroy(ix);
    return fail_cuda(what, e);
  };
  // term_ptr first: its last entry is the number of postings
  if ((e = cudaMalloc(&ix->term_ptr, (static_cast<size_t>(n_terms) + 1) * 8)) != cudaSuccess)
    return bail("cudaMalloc(term_ptr)");
  if ((e = cudaMemcpyAsync(ix->term_ptr, term_ptr, (static_cast<size_t>(n_terms) + 1) * 8,
                           cudaMemcpyDefault, stream)) != cudaSuccess)
    return bail("copy(term_ptr)");
  int64_t nnz = 0;
  if ((e = cudaMemcpyAsync(&nnz, ix->term_ptr + n_terms, 8, cudaMemcpyDeviceToHost, stream)) !=
          cudaSuccess ||
      (e = cudaStreamSynchronize(stream)) != cudaSuccess)
    return bail("read(nnz)");
  if (nnz < 0 || (nnz > 0 && (!post_doc || !post_tf))) {
    cleanup();
    anr_bm25_destroy(ix);
    return fail(ANR_ERR_INVALID, "anr_bm25_create: postings missing or term_ptr corrupt");
  }
  ix->nnz = nnz;
  const size_t pn = static_cast<size_t>(std::max<int64_t>(nnz, 1));
  if ((e = cudaMalloc(&ix->post_doc, pn * 4)) != cudaSuccess) return bail("cudaMalloc(post_doc)");
  if ((e = cudaMalloc(&ix->post_w, pn * 4)) != cudaSuccess) return bail("cudaMalloc(post_w)");
  if ((e = cudaMalloc(&ix->idf, std::max<size_t>(n_terms, 1) * 4)) != cudaSuccess)
    return bail("cudaMalloc(idf)");
  if ((e = cudaMalloc(&tf_dev, pn * 4)) != cudaSuccess) return bail("cudaMalloc(tf)");
  if ((e = cudaMalloc(&dl_dev, std::max<size_t>(n_docs, 1) * 4)) != cudaSuccess)
    return bail("cudaMalloc(doc_len)");
  if ((e = cudaMalloc(&idf64, std::max<size_t>(n_terms, 1) * 8)) != cudaSuccess)
    return bail("cudaMalloc(idf64)");
  if (nnz > 0) {
    if ((e = cudaMemcpyAsync(ix->post_doc, post_doc, pn * 4, cudaMemcpyDefault, stream)) !=
        cudaSuccess)
      return bail("copy(post_doc)");
    if ((e = cudaMemcpyAsync(tf_dev, post_tf, pn * 4, cudaMemcpyDefault, stream)) != cudaSuccess)
      return bail("copy(post_tf)");
  }
  if (n_docs > 0 &&
      (e = cudaMemcpyAsync(dl_dev, doc_len, static_cast<size_t>(n_docs) * 4, cudaMemcpyDefault,
                           stream)) != cudaSuccess)
    return bail("copy(doc_len)");
  if (n_terms > 0) {
    if ((e = cudaMemcpyAsync(idf64, idf, static_cast<size_t>(n_terms) * 8, cudaMemcpyDefault,
                             stream)) != cudaSuccess)
      return bail("copy(idf)");
    f64_to_f32_kernel<<<(n_terms + 255) / 256, 256, 0, stream>>>(idf64, ix->idf, n_terms);
    if ((e = cudaGetLastError()) != cudaSuccess) return bail("idf conversion");
  }
  if ((e = launch_bm25_weights(ix->post_doc, tf_dev, dl_dev, nnz, k1, b, avgdl, ix->post_w,
                               stream)) != cudaSuccess)
    return bail("posting weights");
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return bail("bm25 index build");
  cleanup();
  *out = ix;
  return ANR_OK;
}

int anr_bm25_reweight(anr_ctx* ctx, anr_bm25* index, const int32_t* post_tf, const int32_t* doc_len,
                      const double* idf, double k1, double b, double avgdl) {
  if (!ctx || !index || !idf || (index->nnz > 0 && !post_tf) || (index->n_docs > 0 && !doc_len))
    return fail(ANR_ERR_INVALID, "anr_bm25_reweight: NULL argument");
  if (index->device != ctx->dp.device) return fail(ANR_ERR_INVALID, "index lives on another device");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t pn = static_cast<size_t>(std::max<int64_t>(index->nnz, 1));
  const size_t need = padded(pn * 4) + padded(static_cast<size_t>(std::max(index->n_docs, 1)) * 4) +
                      padded(static_cast<size_t>(std::max(index->n_terms, 1)) * 8) + 2048;
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  const int32_t* tf_dev = post_tf;
  const int32_t* dl_dev = doc_len;
  if (index->nnz > 0 && !is_device_ptr(post_tf)) {
    int32_t* p = arena.take<int32_t>(pn);
    ANR_CUDA(cudaMemcpyAsync(p, post_tf, pn * 4, cudaMemcpyHostToDevice, stream));
    tf_dev = p;
  }
  if (index->n_docs > 0 && !is_device_ptr(doc_len)) {
    int32_t* p = arena.take<int32_t>(static_cast<size_t>(index->n_docs));
    ANR_CUDA(cudaMemcpyAsync(p, doc_len, static_cast<size_t>(index->n_docs) * 4,
                             cudaMemcpyHostToDevice, stream));
    dl_dev = p;
  }
  if (index->n_terms > 0) {
    double* idf64 = arena.take<double>(static_cast<size_t>(index->n_terms));
    ANR_CUDA(cudaMemcpyAsync(idf64, idf, static_cast<size_t>(index->n_terms) * 8, cudaMemcpyDefault,
                             stream));
    f64_to_f32_kernel<<<(index->n_terms + 255) / 256, 256, 0, stream>>>(idf64, index->idf,
                                                                        index->n_terms);
    ANR_CUDA(cudaGetLastError());
  }
  ANR_CUDA(launch_bm25_weights(index->post_doc, tf_dev, dl_dev, index->nnz, k1, b, avgdl,
                               index->post_w, stream));
  // new weights and idf: the dense head rows (and the negative-idf check) are redone by the next search
  if (index->head_built) {
    ANR_CUDA(cudaStreamSynchronize(stream));
    cudaFree(index->head_slot); index->head_slot = nullptr;
    cudaFree(index->head_terms); index->head_terms = nullptr;
    cudaFree(index->head_w); index->head_w = nullptr;
    cudaFree(index->head_max); index->head_max = nullptr;
    index->head_built = false;
    index->n_head = 0;
  }
  index->head_filled = false;
  index->ms_ready = false;   // per-term largest weights and the negative-idf check follow the new values
  ANR_CUDA(cudaStreamSynchronize(stream));   // host sources may be reused by the caller
  return ANR_OK;
}

int anr_bm25_destroy(anr_bm25* index) {
  if (!index) return ANR_OK;
  DeviceGuard guard(index->device);
  if (index->term_ptr) cudaFree(index->term_ptr);
  if (index->post_doc) cudaFree(index->post_doc);
  if (index->post_w) cudaFree(index->post_w);
  if (index->idf) cudaFree(index->idf);
  if (index->head_slot) cudaFree(index->head_slot);
  if (index->head_terms) cudaFree(index->head_terms);
  if (index->head_w) cudaFree(index->head_w);
  if (index->head_max) cudaFree(index->head_max);
  if (index->term_maxw) cudaFree(index->term_maxw);
  if (index->bkt_off) cudaFree(index->bkt_off);
  if (index->bkt_shift) cudaFree(index->bkt_shift);
  if (index->bkt) cudaFree(index->bkt);
  delete index;
  return ANR_OK;
}

int anr_bm25_shape(const anr_bm25* index, int32_t* n_terms, int32_t* n_docs, int64_t* n_postings) {
  if (!index) return fail(ANR_ERR_INVALID, "index is NULL");
  if (n_terms) *n_terms = index->n_terms;
  if (n_docs) *n_docs = index->n_docs;
  if (n_postings) *n_postings = index->nnz;
  return ANR_OK;
}

static int bm25_search_impl(anr_ctx* ctx, const anr_bm25* index, const int32_t* q_terms,
                            const int32_t* q_offsets, int32_t nq, int32_t k,
                            const uint32_t* doc_mask, const int32_t* doc_to_id, int64_t id_base,
                            uint64_t* out_keys, float* out_scores, int32_t* out_docs,
                            int32_t* out_counts, void* stream_v) {
  if (!ctx || !index || !q_offsets) return fail(ANR_ERR_INVALID, "bm25 search: NULL argument");
  if (nq < 1 || k < 1) return fail(ANR_ERR_INVALID, "bm25 search: n_queries and k must be >= 1");
  if (index->device != ctx->dp.device) return fail(ANR_ERR_INVALID, "index lives on another device");
  if (doc_to_id && !is_device_ptr(doc_to_id))
    return fail(ANR_ERR_INVALID, "doc_to_id must be a device pointer (it is index-sized)");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t cells = static_cast<size_t>(nq) * k;
  const size_t need = stage_terms_bytes(q_terms, q_offsets, nq) +
                      stage_mask_bytes(doc_mask, index->n_docs) +
                      bm25_ws_bytes(ctx, index, nq, k) + out_need(out_keys, cells) +
                      out_need(out_scores, cells) + out_need(out_docs, cells) +
                      out_need(out_counts, nq);
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  OutBuf<uint64_t> o_keys = out_make(arena, out_keys, cells);
  OutBuf<float> o_scores = out_make(arena, out_scores, cells);
  OutBuf<int32_t> o_docs = out_make(arena, out_docs, cells);
  OutBuf<int32_t> o_counts = out_make(arena, out_counts, nq);
  TopkOut out;
  out.keys = o_keys.dev;
  out.scores = o_scores.dev;
  out.ids = o_docs.dev;
  out.counts = o_counts.dev;
  out.stride_q = k;
  out.id_base = id_base;
  out.id_map = doc_to_id;
  if (index->n_docs == 0) {
    if (out.keys) ANR_CUDA(cudaMemsetAsync(out.keys, 0, cells * 8, stream));
    if (out.scores) ANR_CUDA(cudaMemsetAsync(out.scores, 0, cells * 4, stream));
    if (out.ids) ANR_CUDA(cudaMemsetAsync(out.ids, 0xff, cells * 4, stream));
    if (out.counts) ANR_CUDA(cudaMemsetAsync(out.counts, 0, static_cast<size_t>(nq) * 4, stream));
  } else {
    QueryTerms qt;
    const uint32_t* mask_dev = nullptr;
    if (int rc = stage_terms(q_terms, q_offsets, nq, arena, stream, &qt)) return rc;
    if (int rc = stage_mask(doc_mask, index->n_docs, arena, stream, &mask_dev)) return rc;
    if (int rc = bm25_pipeline(ctx, index, qt.terms, qt.offsets, nq, k, mask_dev, arena, out,
                               stream))
      return rc;
  }
  bool any_host = false;
  ANR_CUDA(out_flush(o_keys, stream, &any_host));
  ANR_CUDA(out_flush(o_scores, stream, &any_host));
  ANR_CUDA(out_flush(o_docs, stream, &any_host));
  ANR_CUDA(out_flush(o_counts, stream, &any_host));
  if (any_host) ANR_CUDA(cudaStreamSynchronize(stream));
  return ANR_OK;
}

int anr_bm25_search(anr_ctx* ctx, const anr_bm25* index, const int32_t* q_terms,
                    const int32_t* q_offsets, int32_t n_queries, int32_t k,
                    const uint32_t* doc_mask, const int32_t* doc_to_id, int64_t id_base,
                    float* out_scores, int32_t* out_docs, int32_t* out_counts, void* stream) {
  return bm25_search_impl(ctx, index, q_terms, q_offsets, n_queries, k, doc_mask, doc_to_id,
                          id_base, nullptr, out_scores, out_docs, out_counts, stream);
}

int anr_bm25_search_keys(anr_ctx* ctx, const anr_bm25* index, const int32_t* q_terms,
                         const int32_t* q_offsets, int32_t n_queries, int32_t k,
                         const uint32_t* doc_mask, const int32_t* doc_to_id, int64_t id_base,
                         uint64_t* out_keys, void* stream) {
  if (!out_keys) return fail(ANR_ERR_INVALID, "out_keys is NULL");
  return bm25_search_impl(ctx, index, q_terms, q_offsets, n_queries, k, doc_mask, doc_to_id,
                          id_base, out_keys, nullptr, nullptr, nullptr, stream);
}

int anr_bm25_scores(anr_ctx* ctx, const anr_bm25* index, const int32_t* q_terms,
                    int32_t n_q_terms, float* out_scores, void* stream_v) {
  if (!ctx || !index || !out_scores || (n_q_terms > 0 && !q_terms))
    return fail(ANR_ERR_INVALID, "anr_bm25_scores: NULL argument");
  if (n_q_terms < 0) return fail(ANR_ERR_INVALID, "anr_bm25_scores: negative term count");
  if (index->n_docs == 0) return ANR_OK;
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t n = static_cast<size_t>(index->n_docs);
  const size_t need = padded(n * 8) + padded(n * 4) +
                      padded(static_cast<size_t>(n_q_terms + 1) * 4) + 2048;
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  uint64_t* keys = arena.take<uint64_t>(n);
  OutBuf<float> o_scores = out_make(arena, out_scores, n);
  int32_t* off = arena.take<int32_t>(2);
  int32_t* terms = arena.take<int32_t>(static_cast<size_t>(std::max(n_q_terms, 1)));
  set_i32_pair_kernel<<<1, 1, 0, stream>>>(off, 0, n_q_terms);
  ANR_CUDA(cudaGetLastError());
  if (n_q_terms > 0)
    ANR_CUDA(cudaMemcpyAsync(terms, q_terms, static_cast<size_t>(n_q_terms) * 4, cudaMemcpyDefault,
                             stream));
  const Bm25Plan plan = bm25_make_plan(ctx->dp, index->n_docs, 1, 1, true);
  ANR_CUDA(launch_bm25_score_all(bm25_view(index), terms, off, 1, nullptr, plan, keys,
                                 static_cast<int64_t>(n), stream));
  ANR_CUDA(launch_keys_to_scores(keys, static_cast<int64_t>(n), o_scores.dev, stream));
  bool any_host = false;
  ANR_CUDA(out_flush(o_scores, stream, &any_host));
  if (any_host) ANR_CUDA(cudaStreamSynchronize(stream));
  return ANR_OK;
}

// ---- fusion -------------------------------------------------------------------------
int anr_wrrf_fuse(anr_ctx* ctx, const int32_t* ids, const int32_t* lens, const double* weights,
                  int32_t n_lists, int32_t list_stride, int32_t n_queries, double rrf_k,
                  int32_t top_n, int32_t* out_ids, double* out_scores, int32_t* out_counts,
                  void* stream_v) {
  if (!ctx || !ids || !lens || !weights || !out_ids || !out_scores)
    return fail(ANR_ERR_INVALID, "anr_wrrf_fuse: NULL argument");
  if (n_lists < 1 || n_lists > 64 || list_stride < 1 || n_queries < 1 || top_n < 1)
    return fail(ANR_ERR_INVALID, "anr_wrrf_fuse: bad shape");
  if (static_cast<int64_t>(n_lists) * list_stride > (1 << 22))
    return fail(ANR_ERR_UNSUPPORTED, "anr_wrrf_fuse: more than 2^22 entries per query");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t n_ids = static_cast<size_t>(n_queries) * n_lists * list_stride;
  const size_t n_lens = static_cast<size_t>(n_queries) * n_lists;
  const size_t cells = static_cast<size_t>(n_queries) * top_n;
  const bool ids_host = !is_device_ptr(ids), lens_host = !is_device_ptr(lens),
             w_host = !is_device_ptr(weights);
  const size_t scratch_keys = wrrf_scratch_keys(n_lists, list_stride, n_queries);
  const size_t need = (ids_host ? padded(n_ids * 4) + 256 : 0) + padded(scratch_keys * 8) + 256 +
                      (lens_host ? padded(n_lens * 4) + 256 : 0) + (w_host ? 1024 : 0) +
                      out_need(out_ids, cells) + out_need(out_scores, cells) +
                      out_need(out_counts, n_queries);
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  const int32_t* ids_dev = ids;
  const int32_t* lens_dev = lens;
  const double* w_dev = weights;
  if (ids_host) {
    int32_t* p = arena.take<int32_t>(n_ids);
    ANR_CUDA(cudaMemcpyAsync(p, ids, n_ids * 4, cudaMemcpyHostToDevice, stream));
    ids_dev = p;
  }
  if (lens_host) {
    int32_t* p = arena.take<int32_t>(n_lens);
    ANR_CUDA(cudaMemcpyAsync(p, lens, n_lens * 4, cudaMemcpyHostToDevice, stream));
    lens_dev = p;
  }
  if (w_host) {
    double* p = arena.take<double>(n_lists);
    ANR_CUDA(cudaMemcpyAsync(p, weights, static_cast<size_t>(n_lists) * 8, cudaMemcpyHostToDevice,
                             stream));
    w_dev = p;
  }
  OutBuf<int32_t> o_ids = out_make(arena, out_ids, cells);
  OutBuf<double> o_scores = out_make(arena, out_scores, cells);
  OutBuf<int32_t> o_counts = out_make(arena, out_counts, n_queries);
  uint64_t* scratch = scratch_keys ? arena.take<uint64_t>(scratch_keys) : nullptr;
  ANR_CUDA(launch_wrrf_fuse(ids_dev, lens_dev, w_dev, n_lists, list_stride, n_queries, rrf_k, top_n,
                            scratch, o_ids.dev, o_scores.dev, o_counts.dev, stream));
  bool any_host = false;
  ANR_CUDA(out_flush(o_ids, stream, &any_host));
  ANR_CUDA(out_flush(o_scores, stream, &any_host));
  ANR_CUDA(out_flush(o_counts, stream, &any_host));
  if (any_host) ANR_CUDA(cudaStreamSynchronize(stream));
  return ANR_OK;
}

// ---- hybrid ---------------------------------------------------------------------------
int anr_hybrid_search(anr_ctx* ctx, const anr_dense* dense, const anr_bm25* bm25,
                      const float* queries, const int32_t* q_terms, const int32_t* q_offsets,
                      int32_t nq, int32_t k_dense, int32_t k_bm25, const uint32_t* row_mask,
                      const uint32_t* doc_mask, const int32_t* doc_to_id, int64_t id_base,
                      double w_dense, double w_bm25, double rrf_k, int32_t top_n,
                      int32_t* out_ids, double* out_scores, int32_t* out_counts,
                      int32_t* out_dense_rows, float* out_dense_scores, int32_t* out_bm25_ids,
                      float* out_bm25_scores, void* stream_v) {
  if (!ctx || !dense || !bm25 || !queries || !q_offsets || !out_ids || !out_scores)
    return fail(ANR_ERR_INVALID, "anr_hybrid_search: NULL argument");
  if (nq < 1 || k_dense < 1 || k_bm25 < 1 || top_n < 1)
    return fail(ANR_ERR_INVALID, "anr_hybrid_search: bad shape");
  if (dense->device != ctx->dp.device || bm25->device != ctx->dp.device)
    return fail(ANR_ERR_INVALID, "index lives on another device");
  if (dense->n == 0 || bm25->n_docs == 0)
    return fail(ANR_ERR_INVALID, "anr_hybrid_search: empty index");
  if (doc_to_id && !is_device_ptr(doc_to_id))
    return fail(ANR_ERR_INVALID, "doc_to_id must be a device pointer (it is index-sized)");
  const int stride = std::max(k_dense, k_bm25);
  if (2ll * stride > wrrf_max_entries())
    return fail(ANR_ERR_UNSUPPORTED, "anr_hybrid_search: k_dense/k_bm25 above 4096");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);

  const size_t fused_cells = static_cast<size_t>(nq) * top_n;
  const size_t list_cells = static_cast<size_t>(nq) * 2 * stride;
  const size_t need =
      stage_queries_bytes(dense, nq) + stage_mask_bytes(row_mask, dense->n) +
      stage_terms_bytes(q_terms, q_offsets, nq) + stage_mask_bytes(doc_mask, bm25->n_docs) +
      dense_ws_bytes(ctx, dense, nq, k_dense) + bm25_ws_bytes(ctx, bm25, nq, k_bm25) +
      padded(list_cells * 4) + padded(list_cells * 4) + padded(static_cast<size_t>(nq) * 2 * 4) +
      2048 + out_need(out_ids, fused_cells) + out_need(out_scores, fused_cells) +
      out_need(out_counts, nq) + out_need(out_dense_rows, static_cast<size_t>(nq) * k_dense) +
      out_need(out_dense_scores, static_cast<size_t>(nq) * k_dense) +
      out_need(out_bm25_ids, static_cast<size_t>(nq) * k_bm25) +
      out_need(out_bm25_scores, static_cast<size_t>(nq) * k_bm25);
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};

  int32_t* lists = arena.take<int32_t>(list_cells);       // [nq][2][stride] ids
  float* list_scores = arena.take<float>(list_cells);     // [nq][2][stride] scores
  int32_t* lens = arena.take<int32_t>(static_cast<size_t>(nq) * 2);
  double* w_dev = arena.take<double>(2);
  OutBuf<int32_t> o_ids = out_make(arena, out_ids, fused_cells);
  OutBuf<double> o_scores = out_make(arena, out_scores, fused_cells);
  OutBuf<int32_t> o_counts = out_make(arena, out_counts, nq);

  const float* q_dev = nullptr;
  const uint32_t* row_mask_dev = nullptr;
  const uint32_t* doc_mask_dev = nullptr;
  QueryTerms qt;
  if (int rc = stage_queries(dense, queries, nq, dense_padded_queries(ctx, dense, nq, k_dense),
                              arena, stream, &q_dev)) return rc;
  if (int rc = stage_mask(row_mask, dense->n, arena, stream, &row_mask_dev)) return rc;
  if (int rc = stage_terms(q_terms, q_offsets, nq, arena, stream, &qt)) return rc;
  if (int rc = stage_mask(doc_mask, bm25->n_docs, arena, stream, &doc_mask_dev)) return rc;
  const bool fuse_pair = wrrf_fuse_pair_fits(stride);   // weights travel as kernel arguments
  if (!fuse_pair) {
    set_f64_pair_kernel<<<1, 1, 0, stream>>>(w_dev, w_dense, w_bm25);
    ANR_CUDA(cudaGetLastError());
  }

  TopkOut od;
  od.ids = lists;
  od.scores = list_scores;
  od.counts = lens;
  od.stride_q = 2ll * stride;
  od.count_stride = 2;
  od.id_base = id_base;
  TopkOut ob = od;
  ob.ids = lists + stride;
  ob.scores = list_scores + stride;
  ob.counts = lens + 1;
  ob.id_map = doc_to_id;
  // BM25 always runs on the side stream next to the dense pass (ANR_HYBRID_OVERLAP=0 serialises the
  // two).  Default (k_bm25 <= 128): the candidate-driven path -- a chain of short kernels, all of it
  // issued up front -- runs beside the whole dense chain; its kernels ask for the dense kernels' L1
  // split and the GEMM ring leaves them shared memory (kBesideBm25Smem), so the CTAs of both chains
  // truly share the SMs.  With the tiled scan (ANR_BM25_MAXSCORE=0, k_bm25 > 128, negative idf) the
  // older two-phase schedule below applies: sample launch | dense pre-pass | dense main kernel |
  // BM25 main launch, with the GEMM ring capped so that BM25 CTAs of 34 KB fit beside the dense CTA
  // (gemm_ring_cap, anr_dense_gemm.cu).
  static const int overlap_env = getenv("ANR_HYBRID_OVERLAP") ? atoi(getenv("ANR_HYBRID_OVERLAP")) : -1;
  const bool overlap = overlap_env != 0;
  cudaStream_t bm25_stream = overlap ? ctx->side : stream;
  if (ctx->timeline)
    for (bool& b : ctx->tl_set) b = false;
  tl_mark(ctx, 0, stream);
  if (overlap) {
    ANR_CUDA(cudaEventRecord(ctx->ev_fork, stream));
    ANR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  }
  if (overlap) {
    // Tiled scan only: side: BM25 sample launch | main: dense pre-pass kernels, then the dense main
    // kernel | side: the BM25 main launch, held back until the dense main kernel is next in line,
    // so that the persistent, bandwidth-bound dense kernel takes its SMs first (a BM25 main launch
    // that got there first stretched the dense kernel from 0.31 to 0.52 ms) | join.  Gated variant
    // (hybrid_gated, opt-in): ONE BM25 launch held until every dense CTA is resident (DenseGate).
    const bool gated = hybrid_gated(ctx, dense, nq, k_dense);
    DenseGate gate;
    Bm25Run run;
    if (int rc = bm25_pipeline(ctx, bm25, qt.terms, qt.offsets, nq, k_bm25, doc_mask_dev, arena, ob,
                               bm25_stream, !gated, 1, &run))
      return rc;
    tl_mark(ctx, 3, stream);
    // The candidate-driven BM25 path has issued its whole chain in phase 1: nothing is held back,
    // and the dense pass leaves its kernels shared memory; the tiled scan continues in phase 2.
    const bool ms = run.ms_done;
    if (ms) dense_gemm_set_leave_smem(kBesideBm25Smem);
    const int rc_dense = dense_pipeline(ctx, dense, q_dev, nq, k_dense, row_mask_dev, arena, od, stream,
                                        ms ? nullptr : ctx->ev_mid, gated && !ms ? &gate : nullptr);
    dense_gemm_set_leave_smem(0);
    if (rc_dense) return rc_dense;
    tl_mark(ctx, 7, stream);
    if (!ms) ANR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_mid, 0));
    if (gated && !ms) ANR_CUDA(launch_gate_wait(gate.counter, gate.expected, ctx->side));
    if (int rc = bm25_pipeline(ctx, bm25, qt.terms, qt.offsets, nq, k_bm25, doc_mask_dev, arena, ob,
                               bm25_stream, !gated, 2, &run))
      return rc;
    ANR_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
  } else {
    if (int rc = bm25_pipeline(ctx, bm25, qt.terms, qt.offsets, nq, k_bm25, doc_mask_dev, arena, ob,
                               bm25_stream, false))
      return rc;
    if (int rc = dense_pipeline(ctx, dense, q_dev, nq, k_dense, row_mask_dev, arena, od, stream))
      return rc;
  }
  if (overlap) ANR_CUDA(cudaStreamWaitEvent(stream, ctx->ev_join, 0));
  if (fuse_pair)
    ANR_CUDA(launch_wrrf_fuse_pair(lists, lens, w_dense, w_bm25, stride, nq, rrf_k, top_n, o_ids.dev,
                                   o_scores.dev, o_counts.dev, stream));
  else
    ANR_CUDA(launch_wrrf_fuse(lists, lens, w_dev, 2, stride, nq, rrf_k, top_n, nullptr, o_ids.dev,
                              o_scores.dev, o_counts.dev, stream));
  tl_mark(ctx, 11, stream);

  // optional per-retriever outputs: strided device -> user layout [nq, k]
  auto copy_out = [&](void* user, const void* src, int k) -> cudaError_t {
    if (!user) return cudaSuccess;
    return cudaMemcpy2DAsync(user, static_cast<size_t>(k) * 4, src, static_cast<size_t>(stride) * 8,
                             static_cast<size_t>(k) * 4, nq, cudaMemcpyDefault, stream);
  };
  bool any_host = false;
  if (out_dense_rows) any_host |= !is_device_ptr(out_dense_rows);
  if (out_dense_scores) any_host |= !is_device_ptr(out_dense_scores);
  if (out_bm25_ids) any_host |= !is_device_ptr(out_bm25_ids);
  if (out_bm25_scores) any_host |= !is_device_ptr(out_bm25_scores);
  ANR_CUDA(copy_out(out_dense_rows, lists, k_dense));
  ANR_CUDA(copy_out(out_dense_scores, list_scores, k_dense));
  ANR_CUDA(copy_out(out_bm25_ids, lists + stride, k_bm25));
  ANR_CUDA(copy_out(out_bm25_scores, list_scores + stride, k_bm25));
  ANR_CUDA(out_flush(o_ids, stream, &any_host));
  ANR_CUDA(out_flush(o_scores, stream, &any_host));
  ANR_CUDA(out_flush(o_counts, stream, &any_host));
  if (any_host) ANR_CUDA(cudaStreamSynchronize(stream));
  return ANR_OK;
}

// Both local searches of ONE shard in one call, as sortable keys with global ids (what the
// sharded merge consumes): out_keys is [2, n_queries, k], plane 0 = dense, plane 1 = BM25.  Same
// two-stream schedule as anr_hybrid_search (the BM25 chain beside the dense chain); device pointers
// only.
int anr_hybrid_search_keys(anr_ctx* ctx, const anr_dense* dense, const anr_bm25* bm25,
                           const float* queries, const int32_t* q_terms, const int32_t* q_offsets,
                           int32_t nq, int32_t k, const uint32_t* row_mask, const uint32_t* doc_mask,
                           int64_t row_base, int64_t doc_base, uint64_t* out_keys, void* stream_v) {
  if (!ctx || !dense || !bm25 || !queries || !q_offsets || !out_keys)
    return fail(ANR_ERR_INVALID, "anr_hybrid_search_keys: NULL argument");
  if (nq < 1 || k < 1) return fail(ANR_ERR_INVALID, "anr_hybrid_search_keys: bad shape");
  if (dense->device != ctx->dp.device || bm25->device != ctx->dp.device)
    return fail(ANR_ERR_INVALID, "index lives on another device");
  if (dense->n == 0 || bm25->n_docs == 0)
    return fail(ANR_ERR_INVALID, "anr_hybrid_search_keys: empty index");
  if (!is_device_ptr(out_keys) || !is_device_ptr(queries) || !is_device_ptr(q_offsets))
    return fail(ANR_ERR_INVALID, "anr_hybrid_search_keys takes device pointers");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t need = stage_queries_bytes(dense, nq) + stage_mask_bytes(row_mask, dense->n) +
                      stage_terms_bytes(q_terms, q_offsets, nq) +
                      stage_mask_bytes(doc_mask, bm25->n_docs) +
                      dense_ws_bytes(ctx, dense, nq, k) + bm25_ws_bytes(ctx, bm25, nq, k) + 4096;
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  const float* q_dev = nullptr;
  const uint32_t* row_mask_dev = nullptr;
  const uint32_t* doc_mask_dev = nullptr;
  QueryTerms qt;
  if (int rc = stage_queries(dense, queries, nq, dense_padded_queries(ctx, dense, nq, k), arena,
                              stream, &q_dev)) return rc;
  if (int rc = stage_mask(row_mask, dense->n, arena, stream, &row_mask_dev)) return rc;
  if (int rc = stage_terms(q_terms, q_offsets, nq, arena, stream, &qt)) return rc;
  if (int rc = stage_mask(doc_mask, bm25->n_docs, arena, stream, &doc_mask_dev)) return rc;
  TopkOut od;
  od.keys = out_keys;
  od.stride_q = k;
  od.id_base = row_base;
  TopkOut ob;
  ob.keys = out_keys + static_cast<size_t>(nq) * k;
  ob.stride_q = k;
  ob.id_base = doc_base;
  ANR_CUDA(cudaEventRecord(ctx->ev_fork, stream));
  ANR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  const bool gated = hybrid_gated(ctx, dense, nq, k);
  DenseGate gate;
  Bm25Run run;
  if (int rc = bm25_pipeline(ctx, bm25, qt.terms, qt.offsets, nq, k, doc_mask_dev, arena, ob,
                             ctx->side, !gated, 1, &run))
    return rc;
  const bool ms = run.ms_done;   // (see anr_hybrid_search)
  if (ms) dense_gemm_set_leave_smem(kBesideBm25Smem);
  const int rc_dense = dense_pipeline(ctx, dense, q_dev, nq, k, row_mask_dev, arena, od, stream,
                                      ms ? nullptr : ctx->ev_mid, gated && !ms ? &gate : nullptr);
  dense_gemm_set_leave_smem(0);
  if (rc_dense) return rc_dense;
  if (!ms) ANR_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_mid, 0));
  if (gated && !ms) ANR_CUDA(launch_gate_wait(gate.counter, gate.expected, ctx->side));
  if (int rc = bm25_pipeline(ctx, bm25, qt.terms, qt.offsets, nq, k, doc_mask_dev, arena, ob,
                             ctx->side, !gated, 2, &run))
    return rc;
  ANR_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
  ANR_CUDA(cudaStreamWaitEvent(stream, ctx->ev_join, 0));
  return ANR_OK;
}

// ---- sharded merge ----------------------------------------------------------------------
int anr_topk_merge(anr_ctx* ctx, const uint64_t* keys, int32_t n_parts, int32_t n_queries,
                   int32_t k, float* out_scores, int32_t* out_ids, int32_t* out_counts,
                   void* stream_v) {
  if (!ctx || !keys) return fail(ANR_ERR_INVALID, "anr_topk_merge: NULL argument");
  if (n_parts < 1 || n_queries < 1 || k < 1) return fail(ANR_ERR_INVALID, "anr_topk_merge: bad shape");
  if (k > kMaxFusedK) return fail(ANR_ERR_UNSUPPORTED, "anr_topk_merge: k above 128");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t n_keys = static_cast<size_t>(n_parts) * n_queries * k;
  const size_t cells = static_cast<size_t>(n_queries) * k;
  const bool keys_host = !is_device_ptr(keys);
  const size_t need = (keys_host ? padded(n_keys * 8) + 256 : 0) + out_need(out_scores, cells) +
                      out_need(out_ids, cells) + out_need(out_counts, n_queries);
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  const uint64_t* keys_dev = keys;
  if (keys_host) {
    uint64_t* p = arena.take<uint64_t>(n_keys);
    ANR_CUDA(cudaMemcpyAsync(p, keys, n_keys * 8, cudaMemcpyHostToDevice, stream));
    keys_dev = p;
  }
  OutBuf<float> o_scores = out_make(arena, out_scores, cells);
  OutBuf<int32_t> o_ids = out_make(arena, out_ids, cells);
  OutBuf<int32_t> o_counts = out_make(arena, out_counts, n_queries);
  TopkOut out;
  out.scores = o_scores.dev;
  out.ids = o_ids.dev;
  out.counts = o_counts.dev;
  out.stride_q = k;
  ANR_CUDA(launch_topk_final(keys_dev, k, n_parts * k, k, static_cast<int64_t>(n_queries) * k,
                             n_queries, k, out, stream));
  bool any_host = false;
  ANR_CUDA(out_flush(o_scores, stream, &any_host));
  ANR_CUDA(out_flush(o_ids, stream, &any_host));
  ANR_CUDA(out_flush(o_counts, stream, &any_host));
  if (any_host) ANR_CUDA(cudaStreamSynchronize(stream));
  return ANR_OK;
}

int anr_sharded_fuse(anr_ctx* ctx, const uint64_t* gathered, int32_t n_parts, int32_t nq,
                     int32_t k, double w_dense, double w_bm25, double rrf_k, int32_t top_n,
                     int32_t* out_ids, double* out_scores, int32_t* out_counts, void* stream_v) {
  if (!ctx || !gathered || !out_ids || !out_scores)
    return fail(ANR_ERR_INVALID, "anr_sharded_fuse: NULL argument");
  if (n_parts < 1 || nq < 1 || k < 1 || top_n < 1)
    return fail(ANR_ERR_INVALID, "anr_sharded_fuse: bad shape");
  if (k > kMaxFusedK) return fail(ANR_ERR_UNSUPPORTED, "anr_sharded_fuse: k above 128");
  DeviceGuard guard(ctx->dp.device);
  cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : ctx->stream;
  StreamOrder order(ctx, stream);
  const size_t n_keys = static_cast<size_t>(n_parts) * 2 * nq * k;
  const size_t cells = static_cast<size_t>(nq) * top_n;
  const size_t list_cells = static_cast<size_t>(nq) * 2 * k;
  const bool keys_host = !is_device_ptr(gathered);
  const size_t need = (keys_host ? padded(n_keys * 8) + 256 : 0) + padded(list_cells * 4) +
                      padded(static_cast<size_t>(nq) * 2 * 4) + 2048 + out_need(out_ids, cells) +
                      out_need(out_scores, cells) + out_need(out_counts, nq);
  if (int rc = ws_reserve(ctx, need)) return rc;
  Arena arena{ctx->ws, ctx->ws_bytes};
  const uint64_t* keys_dev = gathered;
  if (keys_host) {
    uint64_t* p = arena.take<uint64_t>(n_keys);
    ANR_CUDA(cudaMemcpyAsync(p, gathered, n_keys * 8, cudaMemcpyHostToDevice, stream));
    keys_dev = p;
  }
  int32_t* lists = arena.take<int32_t>(list_cells);  // [nq][2][k]
  int32_t* lens = arena.take<int32_t>(static_cast<size_t>(nq) * 2);
  double* w_dev = arena.take<double>(2);
  OutBuf<int32_t> o_ids = out_make(arena, out_ids, cells);
  OutBuf<double> o_scores = out_make(arena, out_scores, cells);
  OutBuf<int32_t> o_counts = out_make(arena, out_counts, nq);
  static const bool one_launch = !(getenv("ANR_SHARDED_FUSE_SMALL") &&
                                   atoi(getenv("ANR_SHARDED_FUSE_SMALL")) == 0);
  if (one_launch && sharded_fuse_small_fits(n_parts, k)) {   // merge x 2 + fusion in ONE launch
    ANR_CUDA(launch_sharded_fuse_small(keys_dev, n_parts, nq, k, w_dense, w_bm25, rrf_k, top_n,
                                       o_ids.dev, o_scores.dev, o_counts.dev, stream));
    bool any_host_small = false;
    ANR_CUDA(out_flush(o_ids, stream, &any_host_small));
    ANR_CUDA(out_flush(o_scores, stream, &any_host_small));
    ANR_CUDA(out_flush(o_counts, stream, &any_host_small));
    if (any_host_small) ANR_CUDA(cudaStreamSynchronize(stream));
    return ANR_OK;
  }
  set_f64_pair_kernel<<<1, 1, 0, stream>>>(w_dev, w_dense, w_bm25);
  ANR_CUDA(cudaGetLastError());
  const int64_t part_stride = 2ll * nq * k;
  for (int which = 0; which < 2; ++which) {
    TopkOut out;
    out.ids = lists + which * k;
    out.counts = lens + which;
    out.stride_q = 2ll * k;
    out.count_stride = 2;
    ANR_CUDA(launch_topk_final(keys_dev + static_cast<int64_t>(which) * nq * k, k, n_parts * k, k,
                               part_stride, nq, k, out, stream));
  }
  ANR_CUDA(launch_wrrf_fuse(lists, lens, w_dev, 2, k, nq, rrf_k, top_n, nullptr, o_ids.dev,
                            o_scores.dev, o_counts.dev, stream));
  bool any_host = false;
  ANR_CUDA(out_flush(o_ids, stream, &any_host));
  ANR_CUDA(out_flush(o_scores, stream, &any_host));
  ANR_CUDA(out_flush(o_counts, stream, &any_host));
  if (any_host) ANR_CUDA(cudaStreamSynchronize(stream));
  return ANR_OK;
}

}  // extern "C"
