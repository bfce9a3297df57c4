// Internal launcher prototypes shared by the .cu translation units.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace anr {

// Programmatic dependent launch for the kernel chains of a step (anr_common.cuh: pdl_wait /
// pdl_trigger).  The mask (ANR_PDL, default 7; anr_set_option("pdl", mask)) selects who launches
// that way: bit 0 the dense chain, bit 1 the BM25 chain, bit 2 the dense MAIN kernel (made resident
// while the threshold kernel runs); a cleared bit = fully serialised launches, as plain <<<>>>.
constexpr int kPdlDense = 1, kPdlBm25 = 2, kPdlDenseMain = 4;
bool pdl_enabled(int which = kPdlDense);
void pdl_set_mask(int mask);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_on(int which, void (*kern)(KArgs...), dim3 grid, dim3 block,
                                   size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(which) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                cudaStream_t stream, Args... args) {
  return launch_chain_on(kPdlDense, kern, grid, block, smem, stream, args...);
}


// Records the calling thread's anr_last_error() message and returns `code` (anr_api.cu); for the
// translation units that implement ABI entry points of their own.
int set_error(int code, const char* what, const char* detail = nullptr);

struct DeviceProps {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;  // bytes of dynamic shared memory one CTA may opt in to
};

// Where a top-k result goes.  Entry i of query q is written at q * stride_q + i of every
// non-null array; counts[q * count_stride] = number of valid entries.  Ids are
// id_map[id] when id_map is given, else id + id_base.  `keys` receives sortable keys
// with the REMAPPED id (what the sharded merge consumes).
struct TopkOut {
  uint64_t* keys = nullptr;
  float* scores = nullptr;
  int32_t* ids = nullptr;
  int32_t* counts = nullptr;
  int64_t stride_q = 0;
  int64_t count_stride = 1;
  int64_t id_base = 0;
  const int32_t* id_map = nullptr;
  // BM25 only (nullable): CSR offsets of the query terms, q_offsets[q] == q_offsets[q + 1] -> the
  // query has NO result (count 0, empty slots): `if not query_tokens: return []`,
  // src/search_engine.py:216-217
  const int32_t* q_offsets = nullptr;
};

// ---- dense scan -------------------------------------------------------------
// Scans rows [0, n) of emb (row-major, leading dimension ld floats, ld % 4 == 0)
// against nq (1, 2, 4 or 8) queries staged at q_dev ([nq, ld], zero padded) and writes
// per-CTA candidate keys cand[q * cand_stride_q + cta * k + i] (k <= kMaxFusedK).
// Returns the number of CTAs launched through *grid_out (candidates per query = grid * k).
cudaError_t launch_dense_scan_topk(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                   const float* q_dev, int nq, int k, const uint32_t* mask,
                                   uint64_t* cand, int64_t cand_stride_q, int* grid_out,
                                   cudaStream_t stream);
// Upper bound on the grid the scan will use (to size `cand`).
int dense_scan_max_grid(const DeviceProps& dp);
// Largest number of queries (8, 4, 2, 1; 0 = rows too long) one pass can stage in shared
// memory next to a >= 3-stage row ring.
int dense_scan_max_queries(const DeviceProps& dp, int ld, int k, bool emit_all);
// Device-driven exact rescan of flagged queries (see anr_dense.cu): one single-query pass per
// entry of flagged[0 .. *n_flagged), candidates to cand[slot * cand_stride + cta * k + i].
int dense_scan_flagged_grid(const DeviceProps& dp, int64_t n, int ld, int k);
cudaError_t launch_dense_scan_flagged(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                      const float* q_all, const int32_t* n_flagged,
                                      const int32_t* flagged, int k, const uint32_t* mask,
                                      uint64_t* cand, int64_t cand_stride, cudaStream_t stream);
cudaError_t launch_compact_flags(const int32_t* flags, int nq, int32_t* n_flagged, int32_t* flagged,
                                 cudaStream_t stream);
cudaError_t launch_topk_final_flagged(const uint64_t* cand, int64_t cand_stride, int m, int nq,
                                      int k, const TopkOut& out, const int32_t* n_flagged,
                                      const int32_t* flagged, cudaStream_t stream);
// Full materialisation: keys[q * keys_stride_q + row] for every row (0 for masked rows).
cudaError_t launch_dense_scan_all(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                  const float* q_dev, int nq, const uint32_t* mask,
                                  uint64_t* keys, int64_t keys_stride_q, cudaStream_t stream);

// ---- dense scan on the tensor cores (anr_dense_tc.cu) ---------------------------------------
// tcgen05 tf32 scan of 32 queries per pass + exact fp32 rescoring of the nominated candidates.
bool dense_tc_supported(const DeviceProps& dp, int64_t n, int ld, int k);
int dense_tc_queries_per_pass();
size_t dense_tc_cand_keys(const DeviceProps& dp, int64_t n, int k);  // scratch keys one pass needs
cudaError_t launch_row_norm_max(const float* emb, int64_t n, int ld, float* out,
                                cudaStream_t stream);
// ev_start / ev_stop (nullable) are recorded around the main scan kernel only.
// q_dev: [32, ld] (zero rows pad a short group); results of the first n_real queries go through
// `out` (positioned at the group's first query); flags[q] = 1 -> rerun q through the exact scan.
cudaError_t launch_dense_tc(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                            const float* q_dev, int n_real, int k, const uint32_t* mask,
                            float emb_norm_max, uint64_t* cand, const TopkOut& out, int32_t* flags,
                            cudaEvent_t ev_start, cudaEvent_t ev_stop, cudaStream_t stream);

// Same for 64 queries at q_dev ([64, ld]) in ONE pass: a CTA pair shares each corpus tile.
cudaError_t launch_dense_tc_pair(const DeviceProps& dp, const float* emb, int64_t n, int ld,
                                 const float* q_dev, int n_real, int k, const uint32_t* mask,
                                 float emb_norm_max, uint64_t* cand, const TopkOut& out,
                                 int32_t* flags, cudaEvent_t ev_start, cudaEvent_t ev_stop,
                                 cudaStream_t stream);
bool dense_tc_pair_enabled();

// ---- tiled GEMM scan for batches of 64+ queries (anr_dense_gemm.cu) ---------------------------
// 128-row x 64/128/256-query tcgen05 tiles, both operands streamed by TMA; tf32 on the fp32 corpus
// or bf16 on a shadow copy; sample pre-pass thresholds + per-query append buffers + exact fp32
// rescoring (launch_dense_tc_rescore_append).
bool dense_gemm_supported(const DeviceProps& dp, int64_t n, int ld, int k, bool bf16);
int dense_gemm_max_queries();              // queries one launch group takes
int dense_gemm_padded_queries(int nq);     // rows q_dev must hold for a group of nq queries
size_t dense_gemm_scratch_bytes(const DeviceProps& dp, int64_t n, int ld, int nq, int k);
cudaError_t launch_f32_to_bf16(const float* in, void* out, int64_t count, cudaStream_t stream);
// the corpus' bf16 shadow copy (tiled 16 KB blocks, anr_dense_gemm.cu): size and fill
size_t dense_shadow_bytes(int64_t n, int ld);
cudaError_t launch_dense_shadow_fill(const float* emb, void* shadow, int64_t n, int ld,
                                     cudaStream_t stream);
// "The main kernel is resident": every CTA of the persistent main GEMM kernel bumps *counter when
// it starts; expected = its grid.  A hybrid step holds its BM25 launch behind that (a one-thread
// kernel on the BM25 stream polls it, launch_gate_wait), so the dense CTAs own their shared
// memory on every SM BEFORE the BM25 CTAs fill the rest -- an event can only say "the kernel
// before it has finished", which lets a BM25 grid slip in first and starve the dense kernel
// (0.63 instead of 0.41 ms, profiles/r2_call9_*).  counter stays null on paths without one.
struct DenseGate {
  int32_t* counter = nullptr;
  int expected = 0;
};
cudaError_t launch_gate_wait(const int32_t* counter, int expected, cudaStream_t stream);
// GEMM launches issued by the calling thread leave this much shared memory free per SM (0 = none)
void dense_gemm_set_leave_smem(int bytes);
cudaError_t launch_dense_gemm(const DeviceProps& dp, const float* emb, const void* shadow, int64_t n,
                              int ld, const float* q_dev, int n_real, int k, const uint32_t* mask,
                              float emb_norm_max, unsigned char* scratch, const TopkOut& out,
                              int32_t* flags, cudaEvent_t ev_start, cudaEvent_t ev_stop,
                              cudaStream_t stream, cudaEvent_t ev_pre_main = nullptr,
                              DenseGate* gate = nullptr, int32_t* n_flagged = nullptr,
                              int32_t* flagged = nullptr);
cudaError_t launch_dense_tc_rescore_append(const uint64_t* cand, const int32_t* cnt, int cap,
                                           const float* emb, int ld, const float* q_dev, int n_real,
                                           int k, float eps_scale, const uint64_t* thr_key,
                                           const TopkOut& out, int32_t* flags, cudaStream_t stream,
                                           int32_t* ticket = nullptr, int32_t* n_flagged = nullptr,
                                           int32_t* flagged = nullptr);

// ---- top-k ------------------------------------------------------------------
// Per query: select the best k (<= kMaxFusedK) of m candidate keys, sorted best first.
// Candidate i of query q is cand[q * cand_stride_q + (i / seg_len) * seg_stride + i % seg_len].
cudaError_t launch_topk_final(const uint64_t* cand, int64_t cand_stride_q, int m, int seg_len,
                              int64_t seg_stride, int nq, int k, const TopkOut& out,
                              cudaStream_t stream);
// Large-k path: sort keys[q][0..n_pow2) descending in place (global-memory bitonic).
cudaError_t launch_sort_desc(uint64_t* keys, int64_t stride_q, int64_t n_pow2, int nq,
                             cudaStream_t stream);
cudaError_t launch_zero_tail(uint64_t* keys, int64_t stride_q, int64_t n, int64_t n_pow2, int nq,
                             cudaStream_t stream);
cudaError_t launch_emit_sorted(const uint64_t* keys, int64_t stride_q, int64_t n_avail, int nq,
                               int k, const TopkOut& out, cudaStream_t stream);

// ---- BM25 ---------------------------------------------------------------------
struct Bm25View {
  const int64_t* term_ptr;  // [n_terms + 1]
  const int32_t* post_doc;  // [nnz]
  const float* post_w;      // [nnz]
  const float* idf;         // [n_terms]
  int32_t n_terms;
  int32_t n_docs;
  int64_t nnz;
};
// ---- BM25 head terms: dense rows for safe dynamic pruning (anr_bm25.cu) ----------------------
// Terms with df >= n_docs / 8 (at most kBm25MaxHead) are stored a second time as dense rows
// head_w[slot][doc] (0 where the term is absent) so that their weight for ONE document is one
// load.  The top-k kernel then streams only the postings of the other (rare, high-idf) terms,
// bounds what the head terms can still add, and completes the score of the few documents that
// can still reach the tile's k-th best (MaxScore-style pruning; results are unchanged).
constexpr int kBm25MaxHead = 128;
struct Bm25HeadView {
  const uint8_t* slot_of = nullptr;   // [n_terms] head slot of a term, 0xff = not a head term
  const float* head_w = nullptr;      // [n_head][head_ld]
  const float* head_max = nullptr;    // [n_head] largest weight of the row
  int64_t head_ld = 0;
  int32_t n_head = 0;
};
int64_t bm25_head_ld(int n_docs);
// head_max must hold n_head floats; both arrays are (re)written
cudaError_t launch_bm25_head_fill(const Bm25View& ix, const int32_t* head_terms, int n_head,
                                  float* head_w, float* head_max, int64_t head_ld,
                                  cudaStream_t stream);

cudaError_t launch_bm25_weights(const int32_t* post_doc, const int32_t* post_tf,
                                const int32_t* doc_len, int64_t nnz, double k1, double b,
                                double avgdl, float* post_w, cudaStream_t stream);
struct Bm25Plan {
  int tile_docs;   // documents per shared-memory tile (multiple of 32)
  int n_tiles;
  int list_cap;
  int smem_bytes;
  bool beside_dense = false;   // the scan runs on a side stream next to the dense pass (hybrid)
  int phase = 0;               // 0 = whole scan; 1 = only the separate sample launch (if the plan
                               // has one); 2 = everything after it
};
Bm25Plan bm25_make_plan(const DeviceProps& dp, int n_docs, int nq, int k, bool emit_all);
// cand[q * cand_stride_q + tile * k + i] candidate keys (ids = doc index)
// hd (nullable) + theta (nq floats of scratch, nullable) enable the pruned scan.
cudaError_t launch_bm25_score_topk(const Bm25View& ix, const Bm25HeadView* hd, const int32_t* q_terms,
                                   const int32_t* q_offsets, int nq, int k,
                                   const uint32_t* doc_mask, const Bm25Plan& plan, uint64_t* cand,
                                   int64_t cand_stride_q, float* theta, cudaStream_t stream);
// keys[q * keys_stride_q + doc] for every doc (0 for masked docs)
cudaError_t launch_bm25_score_all(const Bm25View& ix, const int32_t* q_terms,
                                  const int32_t* q_offsets, int nq, const uint32_t* doc_mask,
                                  const Bm25Plan& plan, uint64_t* keys, int64_t keys_stride_q,
                                  cudaStream_t stream);
cudaError_t launch_keys_to_scores(const uint64_t* keys, int64_t n, float* scores,
                                  cudaStream_t stream);
// Exhaustive tiled scan of the device-side list of queries q_list[0 .. *n_list): candidates of
// q_list[b] go to cand[b * cand_stride_q + tile * k + i] (the rerun of flagged queries).
cudaError_t launch_bm25_score_listed(const Bm25View& ix, const int32_t* q_terms,
                                     const int32_t* q_offsets, int nq, int k,
                                     const uint32_t* doc_mask, const Bm25Plan& plan, uint64_t* cand,
                                     int64_t cand_stride_q, const int32_t* q_list,
                                     const int32_t* n_list, cudaStream_t stream);

// ---- BM25 top-k, candidate-driven (anr_bm25_ms.cu) ---------------------------------------------
constexpr int kMsMaxTerms = 48;      // query terms the path takes (longer queries are flagged)
constexpr int kMsSample = 4096;      // stage-1 candidates per query
constexpr int kMsSurvivors = 4096;   // stage-2 survivors kept per query (more: flagged)
constexpr int kMsBucketMin = 16;     // posting lists from this length on get a bucket table
// Per-index data of the path (built once, rebuilt after a reweighting): the largest posting weight
// of every term and the bucket tables that make "weight of document d in list t" a two-step lookup.
struct MsIndexView {
  const float* term_maxw = nullptr;    // [n_terms]
  const int64_t* bkt_off = nullptr;    // [n_terms] first entry of the term's table in bkt
  const uint8_t* bkt_shift = nullptr;  // [n_terms] bucket = doc >> shift; 0xff = no table
  const int32_t* bkt = nullptr;        // positions (relative to the list) where the buckets start
};
size_t bm25_ms_scratch_bytes(int nq);
// term_maxw[n_terms] <- largest posting weight of each term; *neg_idf <- 1 if some idf < 0
cudaError_t launch_bm25_term_max(const Bm25View& ix, float* term_maxw, int32_t* neg_idf,
                                 cudaStream_t stream);
int bm25_bucket_shift(int64_t len, int n_docs);        // 0xff: the list gets no table
int64_t bm25_bucket_entries(int shift, int n_docs);    // table entries of a list with that shift
cudaError_t launch_bm25_bucket_fill(const Bm25View& ix, const int64_t* bkt_off, const uint8_t* bkt_shift,
                                    int32_t* bkt, cudaStream_t stream);
// surv: [nq][kMsSurvivors] keys of scratch; the ranked top-k of every query goes to `out`
// (out.q_offsets set: a query without terms gets no result); flagged[0 .. *n_flagged) = the
// queries that must be rerun through the exhaustive scan (ascending), whose rows in `out` are then
// overwritten.  hd.n_head == 0: every lookup goes through the posting lists.
cudaError_t launch_bm25_maxscore(const DeviceProps& dp, const Bm25View& ix, const Bm25HeadView& hd,
                                 const MsIndexView& mx, const int32_t* q_terms,
                                 const int32_t* q_offsets, int nq, int k, const uint32_t* doc_mask,
                                 unsigned char* scratch, uint64_t* surv, const TopkOut& out,
                                 int32_t* n_flagged, int32_t* flagged, cudaStream_t stream,
                                 bool beside_dense /* the kernels share the SMs with a dense pass */,
                                 const cudaEvent_t* marks = nullptr /* [4], nullable entries: recorded
                                 after plan / stage 1 / theta / stage 2 */);

// ---- fusion -------------------------------------------------------------------
// scratch: wrrf_scratch_keys() 64-bit words (0 when the union fits in shared memory)
// two lists (the hybrid step), weights by value; 2 * list_stride <= 64
bool wrrf_fuse_pair_fits(int list_stride);
cudaError_t launch_wrrf_fuse_pair(const int32_t* ids, const int32_t* lens, double w0, double w1,
                                  int list_stride, int nq, double rrf_k, int top_n, int32_t* out_ids,
                                  double* out_scores, int32_t* out_counts, cudaStream_t stream);
cudaError_t launch_wrrf_fuse(const int32_t* ids, const int32_t* lens, const double* weights,
                             int n_lists, int list_stride, int nq, double rrf_k, int top_n,
                             uint64_t* scratch, int32_t* out_ids, double* out_scores,
                             int32_t* out_counts, cudaStream_t stream);
size_t wrrf_scratch_keys(int n_lists, int list_stride, int nq);
// merge of the all-gathered [n_parts][2][nq][k] keys + weighted RRF in one launch (small shapes)
bool sharded_fuse_small_fits(int n_parts, int k);
cudaError_t launch_sharded_fuse_small(const uint64_t* gathered, int n_parts, int nq, int k,
                                      double w_dense, double w_bm25, double rrf_k, int top_n,
                                      int32_t* out_ids, double* out_scores, int32_t* out_counts,
                                      cudaStream_t stream);
int wrrf_max_entries();

}  // namespace anr
