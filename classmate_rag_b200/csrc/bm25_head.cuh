// bm25_head.cuh -- interface between the exact BM25 kernels (bm25.cu) and the batched
// head-matrix path (bm25_mma.cu).
#pragma once
#include "common.cuh"

namespace cmr {

// can this index / call take the head path at all (head_mat + packed postings, no row mask,
// >= 128 documents, tile_docs <= 2048, k + slack <= 128)
bool bm25_head_eligible(const cmr_lex_index& ix, int k, bool has_mask);
size_t bm25_head_workspace_bytes(const cmr_lex_index& ix, int n_queries, int k);
// Writes out_scores / out_ids / out_counts of every certified query and out_flags of all:
// 0 = certified (bit-identical to the exact kernels), != 0 = the query needs the exact kernels.
int bm25_head_topk(const cmr_lex_index& ix, const int* q_terms, const int* q_ptr, int n_queries, int k,
                   long long row_offset, double* out_scores, long long* out_ids, int* out_counts, int* out_flags,
                   void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace cmr
