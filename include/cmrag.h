/*
 * cmrag.h -- C ABI of libcmrag.so, the B200 (sm_100a) implementation of
 * CLASSMATE-RAG's hybrid retrieval hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * reference is pure Python, so the binding a maintainer adds is a ctypes stub
 * (INTEGRATION.md); classmate_rag_b200/_lib.py is that stub.
 *
 * Conventions
 *   - Every pointer marked "device" is CUDA device memory owned by the caller
 *     (the library never allocates or frees caller tensors).  Work is enqueued
 *     on `stream` (a cudaStream_t) and NOT synchronised.
 *   - Row ids handed back are GLOBAL: local row + row_offset of the shard.
 *   - Ranking order everywhere: score descending, then row id ascending
 *     (reference: stable sorted(reverse=True) over insertion order,
 *     rag/retrieval/bm25.py:199; hnswlib ascending distance,
 *     rag/retrieval/vector_chroma.py:204-253).
 *   - Returned scores are float64 and bit-identical to the CPU oracle
 *     (oracle/np_oracle.py): a fast fp32 pass over-selects candidates, an exact
 *     float64 pass in a pinned order rescores them.  out_flags bit 0
 *     (CMR_FLAG_UNCERTIFIED) is set when the fp32 error bound could not prove
 *     the over-selection sufficient; callers then re-run those queries with
 *     cmr_dense_topk_ex(..., CMR_DENSE_EXACT).
 *   - Return value: 0 on success, negative cmr_status otherwise; the message
 *     is available from cmr_last_error() (thread-local).
 */
#ifndef CMRAG_H
#define CMRAG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cmr_stream_t; /* cudaStream_t */

enum cmr_status {
  CMR_OK = 0,
  CMR_EINVAL = -1,       /* bad argument (maps to ValueError in the Python host) */
  CMR_ECUDA = -2,        /* CUDA runtime error (RuntimeError) */
  CMR_EWORKSPACE = -3,   /* workspace too small */
  CMR_EUNSUPPORTED = -4  /* shape outside the compiled kernel set */
};

#define CMR_FLAG_UNCERTIFIED 1
#define CMR_MAX_K 120          /* largest k one pass selects (k + slack <= 128) */
#define CMR_SLACK 8            /* over-selection slack of the fp32 pass */

const char* cmr_last_error(void);
int cmr_version(void);
/* sm count and compute capability of the current device */
int cmr_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------
 * A1  Dense exact top-k.  Replaces the hnswlib search behind
 *     ChromaVectorStore.query (rag/retrieval/vector_chroma.py:204-253).
 *
 *   emb        device, bf16 [n_rows, dim] row-major (dim % 8 == 0, 16 B aligned)
 *   queries    device, bf16 [n_queries, dim]
 *   row_mask   device, uint8 [n_rows] or NULL; 0 = row filtered out (`where`)
 *   out_scores device, float64 [n_queries, k]  exact q.c  (distance = 1 - score)
 *   out_ids    device, int64   [n_queries, k]  global row ids, -1 padded
 *   out_counts device, int32   [n_queries]     valid entries per query
 *   out_flags  device, int32   [n_queries]
 *   cert_eps   absolute bound on |fp32 score - exact score| used by the
 *              certificate (host passes dim * 2^-22 * |q| * max row norm: tensor-pipe
 *              accumulation is not round-to-nearest, so 2^-24 per add is not a bound)
 * ---------------------------------------------------------------------- */
size_t cmr_dense_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k);

int cmr_dense_topk(const uint16_t* emb, int64_t n_rows, int dim,
                   const uint16_t* queries, int n_queries, int k,
                   const uint8_t* row_mask, int64_t row_offset, double cert_eps,
                   double* out_scores, int64_t* out_ids, int32_t* out_counts, int32_t* out_flags,
                   void* workspace, size_t workspace_bytes, cmr_stream_t stream);

/* Same call with an explicit choice of kernel (tests, benchmarks).
 *   CMR_DENSE_SCAN  HBM-streaming scan: coalesced 16-byte loads feed mma.sync tiles, up to
 *                   32 queries per pass over the matrix; warp-private top-k lists.  The
 *                   single-query (GEMV-shaped) path.
 *   CMR_DENSE_MMA   batched q.C^T on the tcgen05 tensor cores: TMA stages 128-byte-swizzled
 *                   tiles of queries and rows in shared memory, one thread issues
 *                   tcgen05.mma (M = 128 queries, N = 256 rows, K = 16) into double-buffered
 *                   TMEM accumulators, four epilogue warps read them back with tcgen05.ld
 *                   (one query per thread) and keep only scores that reach the query's
 *                   admission bound; scores never go to HBM.  The matrix is read once per
 *                   128 queries.  The bound comes from a first pass of the same kernel over
 *                   1/16 of the row tiles (k-th largest of the tile maxima: a valid lower
 *                   bound of the k-th best score), so both passes are exact.
 *                   A row_mask is turned into one bit per row and applied in the epilogue.
 *                   CMR_EUNSUPPORTED when the shape is outside that path (dim < 64, fewer
 *                   than 256 rows, more than 8192 queries).
 *   CMR_DENSE_EXACT exhaustive scan that ranks on the exact float64 dot of EVERY row (fp64 pipe
 *                   bound, a few times slower): nothing to certify, out_flags is always 0.
 *                   The fallback for queries the fast paths flag CMR_FLAG_UNCERTIFIED.
 *   CMR_DENSE_AUTO  SCAN for <= 8 queries, MMA above (when the shape allows).  For a very
 *                   selective mask CMR_DENSE_EXACT is the cheapest: it reads allowed rows only.
 * Results are bit-identical between all of them (ids, order and float64 scores) whenever the
 * fast paths certify theirs. */
#define CMR_DENSE_AUTO 0
#define CMR_DENSE_SCAN 1
#define CMR_DENSE_MMA 2
#define CMR_DENSE_EXACT 3
int cmr_dense_topk_ex(const uint16_t* emb, int64_t n_rows, int dim,
                      const uint16_t* queries, int n_queries, int k,
                      const uint8_t* row_mask, int64_t row_offset, double cert_eps,
                      double* out_scores, int64_t* out_ids, int32_t* out_counts, int32_t* out_flags,
                      void* workspace, size_t workspace_bytes, cmr_stream_t stream, int algo);

/* fp32 -> bf16 (round to nearest even) on device; used for queries and upserts
 * (E5 hands the store fp32, rag/embeddings/__init__.py:85-105). */
int cmr_f32_to_bf16(const float* src, uint16_t* dst, int64_t n, cmr_stream_t stream);

/* ------------------------------------------------------------------------
 * A2  BM25 (Okapi, rank_bm25 semantics) exact top-k.  Replaces
 *     BM25Okapi.get_scores + sorted(...)[:k] inside BM25Store.search
 *     (rag/retrieval/bm25.py:175-212).
 *
 * The inverted index is a term-major CSR over the shard's documents, postings
 * of a term sorted by document, plus a skip table that gives, for the frequent
 * terms, the posting offset at which every tile of `tile_docs` documents starts
 * (rare terms: bisection on the documents of their short lists).
 * Scores are accumulated in float64 in query-token order with rank_bm25's own
 * operation order (idf * factor rounded, then added; no fma) (the per-posting factor is an exact float64, stored either per
 * posting or once per distinct (tf, doc_len) pair), so they are bit-identical to the reference arithmetic and
 * no rescoring pass (and no certificate) is needed: out_flags is always 0.
 * All pointers are device memory.
 * ---------------------------------------------------------------------- */
typedef struct cmr_lex_index {
  const int64_t* term_ptr;   /* [n_terms + 1] posting offsets                          */
  const uint32_t* tile_skip; /* [n_skip_rows, n_tiles + 1] posting offset (relative to term_ptr[t]) at
                                which every tile starts, one row per term that owns one        */
  /* packed postings (preferred, 4 B each): (code << 16) | (doc - tile_lo); code indexes
   * imp_table, the float64 BM25 factor of that posting's (tf, doc_len) pair.            */
  const uint32_t* post_pack; /* [P] or NULL                                             */
  const double* imp_table;   /* [n_codes] or NULL                                       */
  /* wide postings (fallback when the corpus has > 65536 distinct (tf, doc_len) pairs)   */
  const int32_t* post_doc;   /* [P] local document of each posting; may be NULL if packed */
  const double* post_imp;    /* [P] float64 tf*(k1+1)/(tf+k1*(1-b+b*dl/avgdl)); may be NULL if packed */
  const uint16_t* post_tf;   /* [P] term frequency (saturated at 65535); may be NULL    */
  const int32_t* doc_len;    /* [n_docs] tokens per document; may be NULL               */
  const double* idf;         /* [n_terms] float64 idf with the epsilon floor            */
  /* dense terms (optional): terms present in a large share of the shard's documents also
   * get a full per-document column of their float64 factor (0.0 where absent), which the
   * kernel sweeps with coalesced loads instead of scattering their postings.  The CSR
   * arrays still hold their postings (filtered scoring, statistics).                    */
  const double* dense_imp;   /* [n_dense, n_docs] or NULL                               */
  const int32_t* dense_slot; /* [n_terms] column of the term in dense_imp, -1 = sparse; NULL when n_dense == 0 */
  /* head terms (optional; the batched path CMR_BM25_HEAD): the up to 64 terms with the most
   * postings in this shard also get one column each of an fp16 matrix [n_docs, 64] (row pitch
   * 128 bytes, 16-byte aligned) that holds fp16(idf * factor) of the document for the term,
   * 0 where the term is absent.  A batch of queries is scored against it on the tensor cores
   * (the query side is the count of each head term in the query).                        */
  const uint16_t* head_mat;  /* [n_docs, 64] fp16 bits or NULL                          */
  const int32_t* head_slot;  /* [n_terms] column of the term in head_mat, -1 = none; NULL without head_mat */
  /* skip_row [n_terms]: row of the term in tile_skip, or -1: the term (a short list) has no
   * row and its tile slices are found by bisection on post_doc (which must then be present).
   * NULL = tile_skip has a row for every term.  The builder gives rows to the most frequent
   * terms within a memory budget, so the table does not grow with the vocabulary.         */
  const int32_t* skip_row;
  int64_t n_docs;
  int32_t n_terms;
  int32_t tile_docs;         /* multiple of 512, <= 65536                               */
  int32_t n_tiles;
  int32_t n_codes;
  int32_t n_dense;
  int32_t n_head;            /* head_mat columns in use (<= 64)                         */
  double avgdl;
  double k1;
  double b;
} cmr_lex_index;

size_t cmr_bm25_workspace_bytes(const cmr_lex_index* ix, int n_queries, int k);

/*   q_terms    device, int32 [q_ptr[n_queries]] term ids in query-token order,
 *              duplicates kept (rank_bm25 adds a repeated token twice), -1 = unknown
 *   q_ptr      device, int32 [n_queries + 1]
 *   row_mask   device, uint8 [n_docs] or NULL (candidate filter only; subset
 *              statistics are the caller's business)
 *   outputs as cmr_dense_topk; zero-score documents are ranked too, in
 *   ascending id order (bm25.py:199 sorts ALL candidates, stably).
 *   Environment (read per call): CMR_BM25_CTAS_PER_SM=n plans the exact kernel's grid for n
 *   resident CTAs per SM; CMR_BM25_HEAD_MIN_QUERIES=n moves the batch size from which
 *   CMR_BM25_AUTO takes the head-matrix path (default 8).                          */
int cmr_bm25_topk(const cmr_lex_index* ix, const int32_t* q_terms, const int32_t* q_ptr,
                  int n_queries, int k, const uint8_t* row_mask,
                  int64_t row_offset, double* out_scores, int64_t* out_ids,
                  int32_t* out_counts, int32_t* out_flags, void* workspace,
                  size_t workspace_bytes, cmr_stream_t stream);

/* Same call with an explicit choice of kernel (tests, benchmarks).
 *   CMR_BM25_EXACT  bm25_tile_kernel: every (query, tile of documents) accumulates float64 scores
 *                   in shared memory in query-token order -- posting-list scatter for sparse
 *                   terms, column sweep for dense ones -- and keeps the best keys; nothing to
 *                   certify.  Cost grows linearly with the number of queries.
 *   CMR_BM25_HEAD   batched selection on the tensor cores + exact rescoring.  Per block of 32
 *                   queries: (1) the postings of the queries' sparse tokens are bucketed by
 *                   32-document group (CSR scatter in shared memory, one CTA per tile);
 *                   (2) head_mat streams once through TMA into tcgen05.mma (M = 128 documents,
 *                   N = 32 queries, K = 64 head terms, fp32 accumulators in TMEM); the epilogue
 *                   adds each document's bucketed sparse contributions and keeps the documents
 *                   whose fp32 score reaches the query's admission bound (taken from a sample
 *                   pass of the same kernel over 1/16 of the tiles); (3) the KP best candidates
 *                   are rescored exactly -- float64, query-token order, the factor of every
 *                   (token, document) looked up in the posting list -- and ranked.  A query whose
 *                   certificate fails (|fp32 - exact| <= 2^-11-relative bound cannot separate rank
 *                   k from the cut-off; a list or bucket overflowed; more than 16 tokens; a
 *                   negative idf) is re-run by the exact kernel inside the same call, so out_flags
 *                   is still always 0 and the output is bit-identical to CMR_BM25_EXACT.
 *                   CMR_EUNSUPPORTED without head_mat / packed postings, with a row_mask, with
 *                   fewer than 128 documents or tile_docs > 2048.
 *   CMR_BM25_HEAD_NOFALLBACK  as CMR_BM25_HEAD without the exact re-run: out_flags != 0 marks the
 *                   queries that were not certified (their outputs are unspecified).
 *   CMR_BM25_AUTO   HEAD for >= 8 queries when the index and the call allow it, else EXACT.  */
#define CMR_BM25_AUTO 0
#define CMR_BM25_EXACT 1
#define CMR_BM25_HEAD 2
#define CMR_BM25_HEAD_NOFALLBACK 3
int cmr_bm25_topk_ex(const cmr_lex_index* ix, const int32_t* q_terms, const int32_t* q_ptr,
                     int n_queries, int k, const uint8_t* row_mask,
                     int64_t row_offset, double* out_scores, int64_t* out_ids,
                     int32_t* out_counts, int32_t* out_flags, void* workspace,
                     size_t workspace_bytes, cmr_stream_t stream, int algo);

/* ------------------------------------------------------------------------
 * A4  MMR re-ordering of the dense pool (rag/retrieval/fusion.py:39-61,80-102).
 *     cand_rows   device, bf16 [n_queries, pool, dim]: the pool's embeddings
 *                 (cmr_gather_rows fills it from the matrix)
 *     cand_sims   device, float64 [n_queries, pool]: exact q.c (cmr_dense_topk scores)
 *     cand_ids    device, int64 [n_queries, pool];  cand_counts int32 [n_queries]
 *     out_ids / out_sims [n_queries, k] in MMR order, out_counts [n_queries]
 *   Similarities between candidates are the pinned exact float64 dots, the
 *   greedy combine lambda*sim_q - (1-lambda)*max(sim_cc) is float64; first pick
 *   = argmax (lowest index on ties), later picks by strict '>' in ascending index.
 * ---------------------------------------------------------------------- */
int cmr_gather_rows(const uint16_t* emb, int64_t n_rows, int dim, int64_t row_offset,
                    const int64_t* ids, int n_ids, uint16_t* out_rows, cmr_stream_t stream);

int cmr_mmr_select(const uint16_t* cand_rows, const double* cand_sims, const int64_t* cand_ids,
                   const int32_t* cand_counts, int n_queries, int pool, int dim, int k,
                   double lambda, int64_t* out_ids, double* out_sims, int32_t* out_counts,
                   cmr_stream_t stream);

/* ------------------------------------------------------------------------
 * A3 + A5  Reciprocal Rank Fusion, per-id merge and final order of
 *     HybridRetriever.retrieve (rag/retrieval/fusion.py:17-36,108-167).
 *     vec_* : dense list in its final (post-MMR) order, sims = q.c
 *     bm_*  : BM25 list;  kb = 0 gives the non-hybrid path (w_vec is then 1.0)
 *     fused = sum w * (1.0 / (rrf_k + rank)), float64; vector_distance = 1 - sim;
 *     order: (fused, -vector_distance) descending, stable over insertion order
 *     (vector items, then BM25-only items).  out_vdist / out_bm25 are NaN where
 *     the reference has None.
 * ---------------------------------------------------------------------- */
int cmr_hybrid_fuse(const int64_t* vec_ids, const double* vec_sims, const int32_t* vec_counts, int kv,
                    const int64_t* bm_ids, const double* bm_scores, const int32_t* bm_counts, int kb,
                    int n_queries, double w_vec, double w_bm, int rrf_k, int top_k,
                    int64_t* out_ids, double* out_fused, double* out_vdist, double* out_bm25,
                    int32_t* out_counts, cmr_stream_t stream);

/* A3 standalone: rrf_fuse over any number of rank lists (rag/retrieval/fusion.py:17-36).
 *     list_ids [n_lists, max_len] int64 (the caller's integer image of the string ids),
 *     list_counts [n_lists], weights [n_lists] float64 (device).  Output: the distinct ids in
 *     first-appearance order (the reference dict's insertion order) with
 *     score = sum over lists, in list order, of w * (1.0 / (rrf_k + rank)), float64;
 *     out_ids / out_scores hold up to n_lists * max_len entries, out_count [1].
 *     n_lists * max_len <= 4096. */
int cmr_rrf_fuse(const int64_t* list_ids, const int32_t* list_counts, int n_lists, int max_len,
                 const double* weights, int rrf_k, int64_t* out_ids, double* out_scores,
                 int32_t* out_count, cmr_stream_t stream);

/* ------------------------------------------------------------------------
 * N1  Metadata `where` filter on the device (ChromaVectorStore.query's where,
 *     rag/retrieval/vector_chroma.py:45-78,204-253; BM25Store._matches_filter,
 *     rag/retrieval/bm25.py:79-107).  field_codes [n_fields, n_rows] int32: dictionary code
 *     of every row's value per field, -1 = field absent.  The filter is a conjunction of
 *     n_clauses (field, code) equalities; code -1 asks for "absent", a value unknown to the
 *     dictionary is encoded by the host as -2 (matches nothing).  alive (uint8 [n_rows] or
 *     NULL) carries tombstones.  out_mask uint8 [n_rows] feeds row_mask of the top-k calls.
 * ---------------------------------------------------------------------- */
int cmr_filter_mask(const int32_t* field_codes, int64_t n_rows, int n_fields,
                    const int32_t* clause_field, const int32_t* clause_code, int n_clauses,
                    const uint8_t* alive, uint8_t* out_mask, cmr_stream_t stream);

/* N1, statistics of a filtered BM25 search.  The reference rebuilds BM25Okapi over the filtered
 *     entries for every query (rag/retrieval/bm25.py:184-191), so document frequencies are the
 *     subset's.  One pass over the CSR: out_df[t] = number of postings of term t whose document
 *     passes row_mask (uint8 [n_docs], non-zero = passes), out_first[t] = index (into post_doc) of
 *     the first such posting, >= 0x7F7F7F7F when there is none.  n_postings < 2^31 - 1. */
int cmr_masked_df(const int64_t* term_ptr, const int32_t* post_doc, int n_terms, int64_t n_postings,
                  const uint8_t* row_mask, int32_t* out_df, int32_t* out_first, cmr_stream_t stream);

/* ------------------------------------------------------------------------
 * K7  Merge of per-shard top-k lists after the all-gather (multi-GPU; no
 *     reference counterpart).  in_* are [n_parts, n_queries, k]; order is
 *     (score desc, id asc), so the result does not depend on the sharding.
 * ---------------------------------------------------------------------- */
int cmr_topk_merge(const double* in_scores, const int64_t* in_ids, const int32_t* in_counts,
                   int n_parts, int n_queries, int k, double* out_scores, int64_t* out_ids,
                   int32_t* out_counts, cmr_stream_t stream);

/* K7, single exchange: everything a rank contributes to one step -- its dense pool (scores,
 *     ids, flags and the bf16 rows the MMR step reads) and its BM25 list -- packed into one
 *     message per query, so a sharded step has ONE collective: an all-gather of
 *     n_queries * cmr_shard_msg_bytes() bytes per rank.  cmr_shard_merge then merges the
 *     G messages ([n_parts][n_queries][msg]) by (score desc, id asc): dense pool + its rows
 *     ([n_queries, pool, dim], ready for cmr_mmr_select) and the BM25 list.  kb = 0: no
 *     lexical list (non-hybrid); dim = 0: no rows (MMR off). */
size_t cmr_shard_msg_bytes(int pool, int kb, int dim);
int cmr_shard_pack(const double* dense_scores, const int64_t* dense_ids, const int32_t* dense_counts,
                   const int32_t* dense_flags, int pool, const double* bm_scores, const int64_t* bm_ids,
                   const int32_t* bm_counts, int kb, const uint16_t* emb, int64_t n_rows, int dim,
                   int64_t row_offset, int n_queries, void* msg, cmr_stream_t stream);
int cmr_shard_merge(const void* gathered, int n_parts, int n_queries, int pool, int kb, int dim,
                    double* dense_scores, int64_t* dense_ids, int32_t* dense_counts, int32_t* dense_flags,
                    uint16_t* dense_rows, double* bm_scores, int64_t* bm_ids, int32_t* bm_counts,
                    cmr_stream_t stream);

/* K7 over peer memory (no collective launch).  Every rank owns a receive buffer and a few flag
 * words in memory that all ranks of the box have mapped (e.g. a torch symmetric-memory
 * rendezvous: cudaMalloc + IPC / fabric handles over NVLink).  cmr_shard_exchange_pack is
 * cmr_shard_pack with the stores going straight into EVERY rank's receive buffer
 * (slot my_rank of the buffer selected by the step's parity); when its last CTA has finished it
 * raises flag [parity][my_rank] = epoch on every rank (st.release.sys).
 * cmr_shard_exchange_merge waits (ld.acquire.sys) until all n_parts flags of the epoch are
 * up, then merges like cmr_shard_merge.  `state` (two uint32, zero-initialised, rank-local)
 * carries the epoch from step to step, so the pair can be captured in a CUDA graph.  Two
 * buffers alternate (parity): a rank can be at most one step ahead of its slowest peer.
 * A rank that never arrives makes the merge give up after ~10 s and set *timeout_flag. */
typedef struct cmr_shard_p2p {
  const uint64_t* peer_recv;   /* device array [n_parts]: address of every rank's receive buffer   */
  const uint64_t* peer_flags;  /* device array [n_parts]: address of every rank's flag words (u32,
                                  at least 2 * n_parts, zero-initialised)                         */
  uint32_t* state;             /* device, rank-local, 2 x uint32, zero-initialised                */
  int32_t n_parts, my_rank;
  uint64_t slot_stride;        /* bytes reserved per source rank  (>= n_queries * msg bytes)      */
  uint64_t parity_stride;      /* bytes between the two alternating buffers (>= n_parts * slot)   */
  /* optional "pull" form: when every rank's bf16 row matrix is itself mapped into all ranks, the
   * messages carry no rows (they are packed as if dim = 0) and cmr_shard_exchange_merge reads
   * the rows of the merged pool -- pool rows per query instead of n_parts * pool -- straight
   * from their owners over NVLink.  The matrices must not change while a step is in flight. */
  const uint64_t* peer_rows;   /* device array [n_parts]: address of every rank's row matrix, or NULL */
  const int64_t* peer_row_lo;  /* device array [n_parts]: global id of every rank's first row      */
} cmr_shard_p2p;

int cmr_shard_exchange_pack(const double* dense_scores, const int64_t* dense_ids, const int32_t* dense_counts,
                            const int32_t* dense_flags, int pool, const double* bm_scores,
                            const int64_t* bm_ids, const int32_t* bm_counts, int kb, const uint16_t* emb,
                            int64_t n_rows, int dim, int64_t row_offset, int n_queries,
                            const cmr_shard_p2p* x, cmr_stream_t stream);
int cmr_shard_exchange_merge(const void* local_recv, const uint32_t* local_flags, const cmr_shard_p2p* x,
                             int32_t* timeout_flag, int n_queries, int pool, int kb, int dim,
                             double* dense_scores, int64_t* dense_ids, int32_t* dense_counts,
                             int32_t* dense_flags, uint16_t* dense_rows, double* bm_scores, int64_t* bm_ids,
                             int32_t* bm_counts, cmr_stream_t stream);

/* ------------------------------------------------------------------------
 * N3  Query tokeniser on the device (rag/retrieval/bm25.py:34-70,194-195): letter runs of
 *     [A-Za-z] + U+00C0..U+00FF (without U+00D7 / U+00F7), lower-cased, one-character tokens
 *     and the stopwords of the query's language dropped, the rest mapped to term ids (-1 =
 *     not in the vocabulary).  The dictionary is an open-addressing hash table (capacity a
 *     power of two, 64-bit FNV-1a of the lower-cased UTF-8 bytes, slot = (fp ^ fp >> 32) &
 *     (capacity - 1), linear probing, len < 0 = empty slot) over the vocabulary and both
 *     stopword lists; hits are verified byte by byte against `pool`.
 *       flags bit 0: English stopword, bit 1: Italian stopword; val: term id or -1.
 *     text: the queries' UTF-8 bytes back to back, text_ptr int64 [n_queries + 1];
 *     lang_it uint8 [n_queries] (1 = Italian stopword list) or NULL (all English);
 *     out_terms int32 [n_queries, max_terms], padded with -1 (cmr_bm25_topk ignores -1, so
 *     with q_ptr[b] = b * max_terms this is its q_terms input); out_counts int32 [n_queries]
 *     = tokens found (> max_terms means the query was truncated).  All pointers device memory.
 * ---------------------------------------------------------------------- */
typedef struct cmr_token_table {
  const uint64_t* fp;    /* [capacity] */
  const int32_t* off;    /* [capacity] byte offset of the string in pool */
  const int32_t* len;    /* [capacity] byte length, -1 = empty slot */
  const int32_t* val;    /* [capacity] term id or -1 */
  const uint8_t* flags;  /* [capacity] */
  const uint8_t* pool;   /* lower-cased UTF-8 strings */
  int32_t capacity;
  int32_t reserved_;
} cmr_token_table;

int cmr_tokenize_queries(const uint8_t* text, const int64_t* text_ptr, const uint8_t* lang_it,
                         int n_queries, const cmr_token_table* table, int max_terms,
                         int32_t* out_terms, int32_t* out_counts, cmr_stream_t stream);

/* ------------------------------------------------------------------------
 * A9 / K6  Near-duplicate cosine filter over the embedding matrix (extension used by
 *     `rag rebuild`; greedy keep-first rule of rag/utils/dedup.py:40-55: row i is kept iff
 *     no previously KEPT row j < i has q.c >= threshold; hook rag/admin/backup.py:226-233).
 *
 *  cmr_neardup_edges    C . C^T on the tcgen05 tensor cores over the lower triangle (the GEMM
 *                       pipeline of CMR_DENSE_MMA with a threshold epilogue).  Pairs (i, j < i)
 *                       whose fp32 score is >= bound (caller: threshold - error bound) are
 *                       appended to out_edges as (i << 32 | j).  The 128-row blocks
 *                       block_begin, block_begin + block_step, ... are this call's share
 *                       (one call with (0, 1) on a single GPU; rank r of G uses (r, G)).
 *                       out_count may exceed edge_cap: the caller then retries with more room.
 *  cmr_neardup_rescore  exact float64 dot (pinned order) of every candidate pair; pairs with
 *                       exact >= threshold are compacted into out_edges (unordered).
 *  cmr_neardup_resolve  sorted_edges: the surviving keys in ascending order (all shards
 *                       merged).  keep[i] = 1 iff no edge (i, j) has keep[j] = 1.
 * ---------------------------------------------------------------------- */
int cmr_neardup_edges(const uint16_t* emb, int64_t n_rows, int dim, float bound, int block_begin,
                      int block_step, uint64_t* out_edges, uint64_t edge_cap, uint64_t* out_count,
                      cmr_stream_t stream);
int cmr_neardup_rescore(const uint16_t* emb, int dim, const uint64_t* edges, uint64_t n_edges,
                        double threshold, uint64_t* out_edges, uint64_t* out_count, cmr_stream_t stream);
int cmr_neardup_resolve(const uint64_t* sorted_edges, uint64_t n_edges, int64_t n_rows, uint8_t* keep,
                        cmr_stream_t stream);

/* ------------------------------------------------------------------------
 * N4  Ingest-side text dedup: Jaccard similarity between token 5-gram shingle sets
 *     (rag/utils/dedup.py:19-55, called from ingest_file, rag/pipeline/rag.py:308-324).
 *     set_ptr int32 [n_sets + 1], set_items int32: set i = the ascending, unique shingle ids
 *     of chunk i (ids from a host dictionary: equal id <=> equal shingle).  Pairs (i, j < i)
 *     with |A_i ^ A_j| / |A_i v A_j| >= threshold (float64; both empty = 1.0, one empty =
 *     0.0, as _jaccard :32-39) are appended to out_edges as (i << 32 | j), unordered;
 *     out_count may exceed edge_cap (retry with more room).  The greedy keep-first pass of
 *     dedup_text_blocks (:46-53) is cmr_neardup_resolve on the sorted edges.
 * ---------------------------------------------------------------------- */
int cmr_jaccard_edges(const int32_t* set_ptr, const int32_t* set_items, int n_sets, double threshold,
                      uint64_t* out_edges, uint64_t edge_cap, uint64_t* out_count, cmr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CMRAG_H */
