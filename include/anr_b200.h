/* anr_b200.h -- C ABI of the B200-native retrieval hot path for A-NICE-RAG.
 *
 * The reference (pure Python) has no FFI today; these entry points are what a
 * binding for its hot path would call.  Each one names the reference code it
 * replaces (paths relative to the reference repo root).  Conventions:
 *   - every function returns an int status (ANR_OK == 0); no C++ exception ever
 *     crosses the boundary; anr_last_error() returns a thread-local message;
 *   - plain pointers and sizes only.  A data pointer may be a HOST pointer
 *     (pageable or pinned) or a DEVICE pointer on the context's GPU; the library
 *     detects which (cudaPointerGetAttributes).  When any OUTPUT pointer is a
 *     host pointer the call returns after the results have landed; when all
 *     outputs are device pointers the call is asynchronous on `stream`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the context's stream);
 *   - an all-device-pointer search call may be captured into a CUDA graph
 *     (cudaStreamBeginCapture on `stream`) once the same call has run eagerly on
 *     that context: the first call sizes the context's workspace and builds the
 *     lazily created index members, which synchronise.  A captured graph bakes in
 *     workspace addresses, so give it a context nothing else uses
 *     (a-nice-rag_b200/graph.py, tests/test_gpu_graph.py);
 *   - a context (workspace + streams) must not be used by two threads at once; consecutive
 *     calls on one context may use different streams: a call enqueued on another stream than
 *     the previous one is ordered after it (the two share the context's scratch memory);
 *     index objects are immutable after creation and may be shared;
 *   - there is NO CPU implementation behind any of these: without a B200 (sm_100)
 *     device anr_ctx_create fails with ANR_ERR_NO_DEVICE.
 */
#ifndef ANR_B200_H_
#define ANR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANR_ABI_VERSION 1

enum {
  ANR_OK = 0,
  ANR_ERR_INVALID = 1,     /* bad argument */
  ANR_ERR_CUDA = 2,        /* a CUDA call failed; see anr_last_error() */
  ANR_ERR_NO_DEVICE = 3,   /* no sm_100 device / wrong architecture */
  ANR_ERR_OOM = 4,
  ANR_ERR_UNSUPPORTED = 5  /* shape outside what the kernels were built for */
};

typedef struct anr_ctx anr_ctx;
typedef struct anr_dense anr_dense;
typedef struct anr_bm25 anr_bm25;

int anr_abi_version(void);
const char* anr_last_error(void);

/* ---- context ------------------------------------------------------------ */
int anr_ctx_create(int device, anr_ctx** out);
int anr_ctx_destroy(anr_ctx* ctx);
/* Hint: BM25 searches issued through this context run on a stream of their own NEXT TO a dense
 * pass issued through another context (the sharded search, a-nice-rag_b200/sharded.py).  Only the
 * launch shape changes (a separate short sample launch instead of one folded launch). */
int anr_ctx_set_beside_dense(anr_ctx* ctx, int32_t enable);
int anr_ctx_sync(anr_ctx* ctx);
/* sm count, total/free HBM bytes of the context's device (any pointer may be NULL) */
int anr_ctx_info(anr_ctx* ctx, int32_t* sm_count, int64_t* hbm_total, int64_t* hbm_free);

/* How many queries of the LAST search call on this context left the fast paths and were rerun on
 * the device: dense = queries whose tensor-core nomination could not prove exactness (bf16 / tf32
 * margin, overflowing candidate buffers) and went through the exact fp32 scan; bm25 = queries the
 * candidate-driven top-k handed to the exhaustive scan (fewer than k scoring documents, too many
 * survivors, more than 48 terms).  -1 = the call had no such path.  Synchronises the device; valid
 * until the next call on the context.  Both reruns return the same results at a higher cost, so
 * this is the number to watch on production-shaped data. */
int anr_ctx_last_rerun(anr_ctx* ctx, int32_t* dense_queries, int32_t* bm25_queries);

/* Process-wide tuning knobs that can change between calls (measurement scripts; every knob also has
 * an environment variable read at load time).  Known keys:
 *   "pdl"  (ANR_PDL, default 7)  who launches its kernel chain programmatically (each kernel may
 *          become resident while its predecessor runs and waits for it with griddepcontrol.wait):
 *          bit 0 the dense chain, bit 1 the BM25 chain, bit 2 the dense main kernel; 0 = every
 *          launch fully serialised.  No effect on results.
 * Unknown key: ANR_ERR_INVALID.  Not for use while another thread is inside a search call. */
int anr_set_option(const char* key, int32_t value);

/* Per-kernel timing for roofline reports.  While enabled, every launch of the dominant
 * kernels (kind 0 = dense scan kernel -- CUDA-core, tcgen05 or tcgen05 CTA-pair variant --,
 * kind 1 = BM25 score kernel, kind 2 = a whole tensor-core pass: sample pre-pass + threshold +
 * scan + rescoring) is bracketed by CUDA events on the stream it is launched on.  anr_ctx_profile_read synchronises, returns the summed device
 * time (ms) and launch count per kind since the last read, and resets the counters. */
int anr_ctx_profile_enable(anr_ctx* ctx, int32_t on);
int anr_ctx_profile_read(anr_ctx* ctx, int32_t kind, double* total_ms, int64_t* launches);

/* Timeline of ONE anr_hybrid_search step (the GEMM dense path): while enabled (and
 * anr_ctx_profile_enable is off), the call records an event at each of the marks below on the stream
 * that part runs on; anr_ctx_timeline_read synchronises and returns every mark's offset in ms
 * from mark 0 (-1 = not reached).  For naming and timing the fixed costs of a step
 * (bench.py `timeline`); no effect on results.
 *   0 step begin             1 BM25 begin                   2 BM25 candidate-driven top-k end
 *   3 dense pass begin       4 dense main kernel begin      5 dense main kernel end
 *   6 dense rescoring end    7 dense pass end (flagged-query fallback launches included)
 *   8 BM25 rerun begin       9 (same)                      10 BM25 end (rerun of flagged queries)
 *  11 fusion end
 *  12 BM25 plan end         13 BM25 stage 1 end            14 BM25 theta / required lists end
 *  15 BM25 stage 2 end      (2 = the ranking of the survivors' end)
 * (With the tiled BM25 scan, ANR_BM25_MAXSCORE=0: 1-2 = its sample launch, 8-9 = its main launch,
 * 10 = its final top-k; 12-15 are not reached.) */
#define ANR_TIMELINE_MARKS 16
int anr_ctx_timeline_enable(anr_ctx* ctx, int32_t on);
int anr_ctx_timeline_read(anr_ctx* ctx, double* offsets_ms /* [ANR_TIMELINE_MARKS] */);

/* ---- dense index: the chunk-embedding matrix ------------------------------
 * Replaces the per-query `np.stack(df["embedding"].values)` of
 * src/search_engine.py:80,128 with ONE row-major [n, d] fp32 matrix resident in
 * HBM (rows = SQLite `chunks` rows in load order, src/database_manager.py:39-58).
 * emb == NULL allocates an uninitialised index to be filled by
 * anr_dense_upload (streaming loader).  borrow != 0 (device pointers, d % 4 == 0
 * only) makes the index reference the caller's memory instead of copying. */
int anr_dense_create(anr_ctx* ctx, const float* emb, int64_t n, int32_t d, int32_t borrow,
                     anr_dense** out);
int anr_dense_upload(anr_ctx* ctx, anr_dense* index, int64_t row0, const float* rows,
                     int64_t n_rows);
int anr_dense_destroy(anr_dense* index);
int anr_dense_shape(const anr_dense* index, int64_t* n, int32_t* d);
/* enable != 0: batches of more than 32 queries nominate candidates on a bf16 shadow copy of the
 * matrix ([n, d] bf16, built on the next such search; d % 64 == 0) instead of reading the fp32
 * words as tf32: half the HBM bytes and twice the tensor-core rate per pass.  Results do not
 * change: every candidate inside the (wider) error margin is rescored in exact fp32, as in
 * np.dot of src/search_engine.py:81.  Like anr_dense_upload and anr_bm25_reweight this MUTATES the
 * index: do not call it while another thread is searching the same index. */
int anr_dense_set_shadow(anr_dense* index, int32_t enable);
/* The rows of the index have been rewritten behind the library's back (a BORROWED index whose
 * owner updated the tensor in place; anr_dense_upload does this itself for owned indices): drops
 * the cached maximum row norm and the bf16 shadow copy, which the next search rebuilds.  Rows of
 * an index must not change while a search on it is in flight.  The reference re-reads
 * df["embedding"] on every call (src/search_engine.py:80); a resident copy needs this hook. */
int anr_dense_invalidate(anr_dense* index);

/* Inner-product top-k of each query against all (mask-eligible) rows, best
 * first.  Replaces np.dot + argpartition + argsort[::-1] of
 * src/search_engine.py:81-87 (similarity_search_with_embedding) and :129-135
 * (similarity_search); row_mask replaces _filter_by_filename_type (:36-55): bit
 * (i & 31) of word (i >> 5) set = row i eligible; NULL = all rows.
 * Outputs are [n_queries, k]; entries past out_counts[q] = min(k, eligible rows)
 * are score 0 / row -1.  Any k >= 1 is accepted (k >= n gives the full ranking,
 * the `else` branch at :86-87).  Ties: higher score first, then lower row.
 * out_rows holds row + id_base (id_base lets a shard report global rows). */
int anr_dense_search(anr_ctx* ctx, const anr_dense* index, const float* queries,
                     int32_t n_queries, int32_t k, const uint32_t* row_mask, int64_t id_base,
                     float* out_scores, int32_t* out_rows, int32_t* out_counts, void* stream);

/* ---- BM25 index: CSR inverted index --------------------------------------
 * Built from the attributes of the unpickled rank_bm25.BM25Okapi
 * (src/database_manager.py:88-90; built by src/processing/bm25_search.py:77):
 * term t owns postings [term_ptr[t], term_ptr[t+1]) with ascending doc ids,
 * post_tf = doc_freqs[doc][term], doc_len[doc], idf[t] (epsilon floor applied),
 * and k1 / b / avgdl.  The library precomputes the per-posting weight
 *   tf*(k1+1) / (tf + k1*(1 - b + b*doc_len/avgdl))       (float64, stored fp32)
 * which is the bracket of BM25Okapi.get_scores. */
int anr_bm25_create(anr_ctx* ctx, const int64_t* term_ptr, const int32_t* post_doc,
                    const int32_t* post_tf, const int32_t* doc_len, const double* idf,
                    int32_t n_terms, int32_t n_docs, double k1, double b, double avgdl,
                    anr_bm25** out);
/* Re-parameterise an existing index in place: new k1 / b / avgdl (posting weights recomputed
 * from post_tf / doc_len, which the caller supplies again -- the index keeps only the weights)
 * and a new idf table (the epsilon floor is part of it).  The postings themselves are reused:
 * this is the inner step of a k1 / b / epsilon sweep such as src/processing/bm25_test.py:180-315
 * (50 trials over the same corpus).  post_tf is [n_postings] in index order. */
int anr_bm25_reweight(anr_ctx* ctx, anr_bm25* index, const int32_t* post_tf, const int32_t* doc_len,
                      const double* idf, double k1, double b, double avgdl);
int anr_bm25_destroy(anr_bm25* index);
int anr_bm25_shape(const anr_bm25* index, int32_t* n_terms, int32_t* n_docs, int64_t* n_postings);

/* BM25 top-k per query.  Replaces bm25.get_scores(query_tokens) + top-k of
 * src/search_engine.py:219-243 (_core_bm25_search).  Query q is term ids
 * q_terms[q_offsets[q] .. q_offsets[q+1]) in query order, duplicates repeated,
 * -1 = token absent from the vocabulary (contributes 0, like `idf.get(q) or 0`).
 * doc_mask as row_mask above over doc indices (the filtered branch :221-234).
 * doc_to_id (nullable, [n_docs]) maps doc index -> id written to out_docs
 * (the common id space used for fusion); NULL = doc index + id_base.
 * Zero-score documents are legitimate results when fewer than k documents match.
 * A query with NO terms (q_offsets[q] == q_offsets[q+1]) returns nothing (out_counts[q] = 0, keys 0):
 * `if not query_tokens: return []` at src/search_engine.py:216-217; in anr_hybrid_search such a
 * query is fused from its dense list alone. */
int anr_bm25_search(anr_ctx* ctx, const anr_bm25* index, const int32_t* q_terms,
                    const int32_t* q_offsets, int32_t n_queries, int32_t k,
                    const uint32_t* doc_mask, const int32_t* doc_to_id, int64_t id_base,
                    float* out_scores, int32_t* out_docs, int32_t* out_counts, void* stream);

/* Raw scores of every document for ONE query: BM25Okapi.get_scores(query_tokens) itself
 * (src/search_engine.py:219), fp32, out_scores is [n_docs]. */
int anr_bm25_scores(anr_ctx* ctx, const anr_bm25* index, const int32_t* q_terms,
                    int32_t n_q_terms, float* out_scores, void* stream);

/* ---- weighted reciprocal-rank fusion -------------------------------------
 * Replaces SearchEngine.weighted_reciprocal_rank_fusion (src/search_engine.py:21-34):
 * score[id] += weights[l] * (1 / (rrf_k + rank)), rank from 1, lists in order,
 * float64 with the reference's operation order (bit-identical scores); output
 * sorted by score descending, ties in first-insertion order (list 0 first).
 * ids is [n_queries, n_lists, list_stride], lens is [n_queries, n_lists].
 * Writes the first min(top_n, |union|) fused entries per query.  Up to 2^22 entries per query
 * (n_lists * list_stride); unions above 8192 entries use an L2-resident scratch. */
int anr_wrrf_fuse(anr_ctx* ctx, const int32_t* ids, const int32_t* lens, const double* weights,
                  int32_t n_lists, int32_t list_stride, int32_t n_queries, double rrf_k,
                  int32_t top_n, int32_t* out_ids, double* out_scores, int32_t* out_counts,
                  void* stream);

/* ---- the whole hybrid query in one call -----------------------------------
 * dense top-k_dense + BM25 top-k_bm25 + WRRF -> top_n, i.e. the body of
 * RetrievalEvaluationSystem.retrieve_documents (src/query_rag_retrieval.py:206-212,
 * :308-315, :357-362) for one dense model + BM25, for n_queries queries at once.
 * Dense ids are rows (+ id_base); BM25 ids go through doc_to_id (NULL = doc +
 * id_base).  List order for tie purposes: dense first, then BM25.
 * out_dense_rows / out_dense_scores / out_bm25_ids / out_bm25_scores are optional
 * ([n_queries, k_*], NULL to skip). */
int anr_hybrid_search(anr_ctx* ctx, const anr_dense* dense, const anr_bm25* bm25,
                      const float* queries, const int32_t* q_terms, const int32_t* q_offsets,
                      int32_t n_queries, int32_t k_dense, int32_t k_bm25,
                      const uint32_t* row_mask, const uint32_t* doc_mask,
                      const int32_t* doc_to_id, int64_t id_base, double w_dense, double w_bm25,
                      double rrf_k, int32_t top_n, int32_t* out_ids, double* out_scores,
                      int32_t* out_counts, int32_t* out_dense_rows, float* out_dense_scores,
                      int32_t* out_bm25_ids, float* out_bm25_scores, void* stream);

/* ---- corpus sharding (one process per GPU) --------------------------------
 * Local top-k as sortable 64-bit keys (score in the high word, id in the low
 * word) so that an all-gather of [n_queries, k] keys per rank followed by
 * anr_topk_merge gives the same result as an unsharded search.  keys are
 * [n_queries, k], best first, 0 = empty slot. */
int anr_dense_search_keys(anr_ctx* ctx, const anr_dense* index, const float* queries,
                          int32_t n_queries, int32_t k, const uint32_t* row_mask,
                          int64_t id_base, uint64_t* out_keys, void* stream);
int anr_bm25_search_keys(anr_ctx* ctx, const anr_bm25* index, const int32_t* q_terms,
                         const int32_t* q_offsets, int32_t n_queries, int32_t k,
                         const uint32_t* doc_mask, const int32_t* doc_to_id, int64_t id_base,
                         uint64_t* out_keys, void* stream);
/* Both local searches of one SHARD in one call (a-nice-rag_b200/sharded.py): out_keys is
 * [2, n_queries, k] sortable keys with global ids (plane 0 = dense rows + row_base, plane 1 = BM25
 * documents + doc_base), the input of anr_sharded_fuse after the all-gather.  Device pointers
 * only; BM25 runs on the context's side stream around the dense pass as in anr_hybrid_search. */
int anr_hybrid_search_keys(anr_ctx* ctx, const anr_dense* dense, const anr_bm25* bm25,
                           const float* queries, const int32_t* q_terms, const int32_t* q_offsets,
                           int32_t n_queries, int32_t k, const uint32_t* row_mask,
                           const uint32_t* doc_mask, int64_t row_base, int64_t doc_base,
                           uint64_t* out_keys, void* stream);

/* keys: [n_parts, n_queries, k] (the all-gather output) -> merged top-k. */
int anr_topk_merge(anr_ctx* ctx, const uint64_t* keys, int32_t n_parts, int32_t n_queries,
                   int32_t k, float* out_scores, int32_t* out_ids, int32_t* out_counts,
                   void* stream);

/* The consumer of the all-gathered buffer of a sharded hybrid query: gathered is
 * [n_parts][2][n_queries][k] keys (per rank: dense keys, then BM25 keys, ids already global).
 * Merges each retriever's n_parts lists to its global top-k and runs the weighted RRF on the
 * two merged lists (fusion needs GLOBAL ranks, so it follows the merge) -> top_n fused.
 * Same outputs as anr_hybrid_search. */
int anr_sharded_fuse(anr_ctx* ctx, const uint64_t* gathered, int32_t n_parts, int32_t n_queries,
                     int32_t k, double w_dense, double w_bm25, double rrf_k, int32_t top_n,
                     int32_t* out_ids, double* out_scores, int32_t* out_counts, void* stream);

/* ---- bulk loader (host only, no device involved) ------------------------------------------
 * Replaces the per-row Python loop of DatabaseManager.load_embeddings_from_sql
 * (src/database_manager.py:35-63: fetchall + np.frombuffer per row): `sql` must yield
 * (rowid, blob) per row, e.g. "SELECT rowid, embedding FROM chunks"; every BLOB of exactly
 * row_bytes is copied to dst + i * row_bytes (dst: host memory, pinned or pageable, room for
 * max_rows rows) and its rowid to rowids[i] (may be NULL).  The scan stops at the first row that
 * is not a BLOB of that size (NULL, ragged width -- the reference skips or keeps such rows one by
 * one) or at row max_rows: *uniform = 0 then, and the caller falls back to its row-by-row path;
 * *uniform = 1 means the whole table was copied, *n_rows rows.  libsqlite3.so.0 is resolved at
 * run time (ANR_ERR_UNSUPPORTED when it is absent). */
int anr_sqlite_read_blobs(const char* db_path, const char* sql, void* dst, int64_t row_bytes,
                          int64_t max_rows, int64_t* rowids, int64_t* n_rows, int32_t* uniform);

#ifdef __cplusplus
}
#endif
#endif /* ANR_B200_H_ */
