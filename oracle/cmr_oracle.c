/* cmr_oracle.c -- plain C restatement of the two numeric cores of the hybrid
 * retrieval path, for parity checks at sizes where NumPy is slow.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/, smoke() and
 * bench.py's cpu_baseline leg may load this; the product never does.
 *
 *  - oracle_exact_dots: float64 dot of a bf16 query with bf16 rows in the pinned
 *    order documented in oracle/np_oracle.py (replaces the hnswlib cosine search
 *    behind ChromaVectorStore.query, rag/retrieval/vector_chroma.py:204-253).
 *  - oracle_bm25_scores: rank_bm25.BM25Okapi.get_scores over a CSR index,
 *    float64, operation for operation (call site rag/retrieval/bm25.py:197;
 *    rank_bm25 itself is not vendored: parity unpinned for it).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: every
 * operation must round exactly like NumPy's).
 */
#include <stdint.h>
#include <string.h>

static double bf16_to_f64(uint16_t b) {
  uint32_t u = ((uint32_t)b) << 16;
  float f;
  memcpy(&f, &u, 4);
  return (double)f;
}

void oracle_exact_dots(const uint16_t* q, const uint16_t* rows, int64_t n_rows, int dim, double* out) {
  const int nvec = dim / 8;
  for (int64_t r = 0; r < n_rows; ++r) {
    const uint16_t* c = rows + r * (int64_t)dim;
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0;
    for (int v = 0; v < nvec; ++v) {         /* vector v belongs to lane v % 32, visited in order */
      const int l = v & 31;
      for (int e = 0; e < 8; ++e) {
        const int i = 8 * v + e;
        acc[l] = acc[l] + bf16_to_f64(q[i]) * bf16_to_f64(c[i]);
      }
    }
    for (int off = 16; off >= 1; off >>= 1)
      for (int l = 0; l < off; ++l) acc[l] = acc[l] + acc[l + off];
    out[r] = acc[0];
  }
}

void oracle_bm25_scores(const int64_t* term_ptr, const int32_t* post_doc, const int32_t* post_tf,
                        const int32_t* doc_len, const double* idf, double avgdl, double k1, double b,
                        const int32_t* q_terms, int n_q_terms, int32_t n_terms, int64_t n_docs, double* score) {
  for (int64_t d = 0; d < n_docs; ++d) score[d] = 0.0;
  const double one_minus_b = 1.0 - b;
  const double k1p1 = k1 + 1.0;
  for (int j = 0; j < n_q_terms; ++j) {
    const int32_t t = q_terms[j];
    if (t < 0 || t >= n_terms) continue;
    for (int64_t p = term_ptr[t]; p < term_ptr[t + 1]; ++p) {
      const int32_t d = post_doc[p];
      const double tf = (double)post_tf[p];
      const double num = tf * k1p1;
      const double norm = one_minus_b + (b * (double)doc_len[d]) / avgdl;
      const double den = tf + k1 * norm;
      score[d] = score[d] + idf[t] * (num / den);
    }
  }
}
