// lex_slice.cuh -- where a term's postings of one tile of documents are.
//
// Postings of a term are sorted by document.  The most frequent terms (as many as a memory budget
// allows, chosen at build time) own a row of the skip table: the posting offset at which every
// tile starts, so a slice is two loads.  Every other term (skip_row[t] < 0) has a short list:
// its slice is found by bisection on the documents (post_doc).  A vocabulary of a million terms
// therefore costs 4 bytes per term, not a row of n_tiles + 1 offsets each.
#pragma once
#include "common.cuh"

namespace cmr {

// first posting index in [lo, hi) whose document is >= doc
__device__ __forceinline__ long long lex_lower_bound(const int* __restrict__ post_doc, long long lo, long long hi,
                                                     long long doc) {
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)post_doc[mid] < doc) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// row of term t in the skip table, or -1
__device__ __forceinline__ int lex_skip_row(const cmr_lex_index& ix, int t) {
  return ix.skip_row != nullptr ? ix.skip_row[t] : t;   // no map: the table has a row per term
}

// [*lo, *hi): offsets RELATIVE to term_ptr[t] of the postings of term t that fall into `tile`
__device__ __forceinline__ void lex_slice(const cmr_lex_index& ix, int t, int row, int tile, u32* lo, u32* hi) {
  if (row >= 0) {
    const u32* sk = ix.tile_skip + (size_t)row * (ix.n_tiles + 1) + tile;
    *lo = sk[0];
    *hi = sk[1];
    return;
  }
  const long long base = ix.term_ptr[t], end = ix.term_ptr[t + 1];
  const long long d_lo = (long long)tile * ix.tile_docs;
  const long long a = lex_lower_bound(ix.post_doc, base, end, d_lo);
  const long long z = lex_lower_bound(ix.post_doc, a, end, d_lo + ix.tile_docs);
  *lo = (u32)(a - base);
  *hi = (u32)(z - base);
}

}  // namespace cmr
