// text.cu -- N3: query tokeniser on the device (reference rag/retrieval/bm25.py:34-70,194-195).
//
// Batches of thousands of queries (config C4: 4096 per batch) would otherwise be tokenised
// by a Python loop on the host.  One thread walks one query's UTF-8 bytes and reproduces
// `_tokenize` exactly:
//   * a token is a maximal run of letters of [A-Za-z] and U+00C0..U+00FF without U+00D7 (x)
//     and U+00F7 (/) -- in UTF-8: an ASCII letter, or 0xC3 followed by 0x80..0xBF except
//     0x97 and 0xB7; every other byte sequence separates tokens;
//   * str.lower(): A-Z -> a-z, U+00C0..U+00DE -> +0x20 (0xC3 0x80..0x9E -> second byte +0x20);
//   * tokens of one character are dropped (characters, not bytes);
//   * tokens in the stopword set of the query's language (English unless the tag starts
//     with "it") are dropped;
//   * the rest is looked up in the vocabulary: term id, or -1 when unseen (an unseen token
//     contributes nothing to BM25 but is kept, as in the reference).
// The dictionary is one open-addressing hash table over the lower-cased UTF-8 strings of the
// vocabulary and of both stopword lists (64-bit FNV-1a, verified byte by byte on a hit).
// Output: term ids [n_queries, max_terms] padded with -1 (cmr_bm25_topk ignores -1 tokens, so
// the padded rows are directly a CSR with q_ptr[b] = b * max_terms) and the true token counts.
#include "common.cuh"

namespace cmr {

__device__ __forceinline__ bool is_ascii_letter(unsigned c) { return (c | 0x20u) - 'a' < 26u; }

// lower-cased byte at position i of the token starting at s (the previous byte decides
// whether this is the second byte of a 0xC3 pair)
__device__ __forceinline__ unsigned lower_byte(const uint8_t* s, int i) {
  const unsigned c = s[i];
  if (c < 0x80u) return (c - 'A' < 26u) ? c + 0x20u : c;
  if (c != 0xC3u && i > 0 && s[i - 1] == 0xC3u && c >= 0x80u && c <= 0x9Eu) return c + 0x20u;
  return c;
}

__device__ __forceinline__ int table_lookup(const cmr_token_table& t, const uint8_t* tok, int n_bytes,
                                            unsigned long long fp, unsigned* flags) {
  unsigned slot = (unsigned)(fp ^ (fp >> 32)) & (unsigned)(t.capacity - 1);
  for (int probe = 0; probe < t.capacity; ++probe) {
    const int len = t.len[slot];
    if (len < 0) break;  // empty slot: not in the dictionary
    if (t.fp[slot] == fp && len == n_bytes) {
      const uint8_t* ref = t.pool + t.off[slot];
      bool same = true;
      for (int i = 0; i < n_bytes && same; ++i) same = lower_byte(tok, i) == ref[i];
      if (same) {
        *flags = t.flags[slot];
        return t.val[slot];
      }
    }
    slot = (slot + 1) & (unsigned)(t.capacity - 1);
  }
  *flags = 0;
  return -1;
}

__global__ void tokenize_kernel(const uint8_t* __restrict__ text, const long long* __restrict__ text_ptr,
                                const uint8_t* __restrict__ lang_it, int n_queries, cmr_token_table tab,
                                int max_terms, int* __restrict__ out_terms, int* __restrict__ out_counts) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  const uint8_t* s = text + text_ptr[q];
  const int n = (int)(text_ptr[q + 1] - text_ptr[q]);
  const unsigned stop_bit = (lang_it != nullptr && lang_it[q]) ? 2u : 1u;
  int* dst = out_terms + (size_t)q * max_terms;
  int count = 0;
  int i = 0;
  while (i < n) {
    // find the next token start
    int start = -1;
    while (i < n) {
      const unsigned c = s[i];
      if (is_ascii_letter(c)) { start = i; break; }
      if (c == 0xC3u && i + 1 < n) {
        const unsigned d = s[i + 1];
        if (d >= 0x80u && d <= 0xBFu && d != 0x97u && d != 0xB7u) { start = i; break; }
        i += 2;
        continue;
      }
      ++i;
    }
    if (start < 0) break;
    // walk the run
    unsigned long long fp = 14695981039346656037ull;
    int chars = 0;
    while (i < n) {
      const unsigned c = s[i];
      if (is_ascii_letter(c)) {
        fp = (fp ^ ((c - 'A' < 26u) ? c + 0x20u : c)) * 1099511628211ull;
        ++i;
        ++chars;
      } else if (c == 0xC3u && i + 1 < n) {
        const unsigned d = s[i + 1];
        if (!(d >= 0x80u && d <= 0xBFu && d != 0x97u && d != 0xB7u)) break;
        fp = (fp ^ 0xC3u) * 1099511628211ull;
        fp = (fp ^ ((d <= 0x9Eu) ? d + 0x20u : d)) * 1099511628211ull;
        i += 2;
        ++chars;
      } else {
        break;
      }
    }
    if (chars > 1) {
      unsigned flags = 0;
      const int id = table_lookup(tab, s + start, i - start, fp, &flags);
      if (!(flags & stop_bit)) {
        if (count < max_terms) dst[count] = id;
        ++count;
      }
    }
  }
  for (int j = count; j < max_terms; ++j) dst[j] = -1;
  out_counts[q] = count;
}


// ---------------------------------------------------------------------------------------
// N4  Jaccard near-duplicate edges between shingle sets (ingest-side text dedup,
//     rag/utils/dedup.py:19-55).  Set i = the sorted, unique ids of chunk i's token 5-gram
//     shingles (the host maps shingle tuples to ids through a dictionary, so equal ids <=>
//     equal shingles: no hashing, no collisions).  An edge (i, j < i) is emitted when
//         |A_i ^ A_j| / |A_i v A_j| >= threshold      (float64 division, like Python's)
//     with the reference's conventions: two empty sets 1.0, one empty set 0.0.
//     One CTA per set i (largest first: the triangle's long rows start early), set i staged in
//     shared memory; warp w takes j = w, w + 8, ...: the lanes read set j coalesced and bisect
//     set i.  A pair whose sizes already rule the threshold out (min/max < threshold; the
//     division is monotonic, so this is exact) is skipped without touching the sets.
// ---------------------------------------------------------------------------------------
constexpr int JAC_THREADS = 256;
constexpr int JAC_SMEM_ITEMS = 8192;  // a larger set i is bisected in global memory

__global__ void __launch_bounds__(JAC_THREADS)
jaccard_edges_kernel(const int* __restrict__ set_ptr, const int* __restrict__ items, int n_sets, double threshold,
                     unsigned long long* __restrict__ edges, unsigned long long edge_cap,
                     unsigned long long* __restrict__ edge_count) {
  __shared__ int s_a[JAC_SMEM_ITEMS];
  const int i = n_sets - 1 - (int)blockIdx.x;
  const int a_lo = set_ptr[i], na = set_ptr[i + 1] - a_lo;
  const int* a = items + a_lo;
  if (na <= JAC_SMEM_ITEMS) {
    for (int t = threadIdx.x; t < na; t += JAC_THREADS) s_a[t] = a[t];
    __syncthreads();
    a = s_a;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < i; j += JAC_THREADS / 32) {
    const int b_lo = set_ptr[j], nb = set_ptr[j + 1] - b_lo;
    double jac;
    if (na == 0 && nb == 0) {
      jac = 1.0;
    } else if (na == 0 || nb == 0) {
      jac = 0.0;
    } else {
      const int mn = na < nb ? na : nb, mx = na < nb ? nb : na;
      if ((double)mn / (double)mx < threshold) continue;  // |A^B| <= mn and |AvB| >= mx
      const int* b = items + b_lo;
      int inter = 0;
      for (int t = lane; t < nb; t += 32) {
        const int x = b[t];
        int lo = 0, hi = na;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (a[mid] < x) lo = mid + 1;
          else hi = mid;
        }
        inter += (lo < na && a[lo] == x) ? 1 : 0;
      }
      inter = __reduce_add_sync(0xFFFFFFFFu, inter);
      jac = (double)inter / (double)(na + nb - inter);
    }
    if (lane == 0 && jac >= threshold) {
      const unsigned long long slot = atomicAdd(edge_count, 1ull);
      if (slot < edge_cap) edges[slot] = ((unsigned long long)(unsigned)i << 32) | (unsigned long long)(unsigned)j;
    }
  }
}

}  // namespace cmr

using namespace cmr;

extern "C" int cmr_tokenize_queries(const uint8_t* text, const int64_t* text_ptr, const uint8_t* lang_it,
                                    int n_queries, const cmr_token_table* table, int max_terms,
                                    int32_t* out_terms, int32_t* out_counts, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_queries >= 0 && max_terms >= 1, "bad tokenizer shape");
  if (n_queries == 0) return CMR_OK;
  CMR_CHECK_ARG(text_ptr && out_terms && out_counts && table, "null pointer argument");
  CMR_CHECK_ARG(table->capacity >= 2 && (table->capacity & (table->capacity - 1)) == 0 && table->fp && table->off &&
                    table->len && table->val && table->flags && table->pool,
                "token table must have a power-of-two capacity and all arrays");
  tokenize_kernel<<<(n_queries + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      text, (const long long*)text_ptr, lang_it, n_queries, *table, max_terms, out_terms, out_counts);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_jaccard_edges(const int32_t* set_ptr, const int32_t* set_items, int n_sets, double threshold,
                                 uint64_t* out_edges, uint64_t edge_cap, uint64_t* out_count, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_sets >= 0, "n_sets negative");
  CMR_CHECK_ARG(out_count != nullptr, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  CMR_CUDA(cudaMemsetAsync(out_count, 0, sizeof(uint64_t), st));
  if (n_sets < 2) return CMR_OK;
  CMR_CHECK_ARG(set_ptr && (out_edges || edge_cap == 0), "null pointer argument");
  jaccard_edges_kernel<<<n_sets, JAC_THREADS, 0, st>>>(set_ptr, set_items, n_sets, threshold,
                                                      (unsigned long long*)out_edges, edge_cap,
                                                      (unsigned long long*)out_count);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}
