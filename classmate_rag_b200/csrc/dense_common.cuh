// dense_common.cuh -- pieces shared by the two dense top-k paths (dense.cu: HBM-streaming
// scan on mma.sync; dense_mma.cu: batched tcgen05/TMA GEMM).  Both end the same way: the KP
// best fp32 keys of a query are rescored exactly in float64 and ordered.
#pragma once
#include "topk.cuh"

namespace cmr {

constexpr int FIN_THREADS = 1024;

// Tail of both finalize kernels.  s_out[KP]: selected keys, sorted best first, empties (0)
// last.  One warp per candidate recomputes the exact float64 dot (pinned order, bit-identical
// to oracle/np_oracle.py:exact_dots); the final order is (exact score desc, row asc).
// `extra_flag` is OR-ed into the query's flag (candidate-buffer overflow of the GEMM path).
// Must be called by all FIN_THREADS threads; s_score[KP] is scratch.
template <int KP>
__device__ __forceinline__ void dense_finalize_tail(const u64* s_out, double* s_score,
                                                    const uint16_t* __restrict__ emb, int dim,
                                                    const uint16_t* __restrict__ q, long long row_offset,
                                                    int k, double cert_eps, int extra_flag, int qi,
                                                    double* __restrict__ out_scores,
                                                    long long* __restrict__ out_ids,
                                                    int* __restrict__ out_counts, int* __restrict__ out_flags) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = warp; c < KP; c += FIN_THREADS / 32) {
    const u64 key = s_out[c];
    if (key == 0ull) continue;  // warp-uniform
    const uint16_t* row = emb + (size_t)key_row(key) * dim;
    const double s = warp_exact_dot(q, row, dim, lane);
    if (lane == 0) s_score[c] = s;
  }
  __syncthreads();

  const int n_valid = count_valid(s_out, KP);  // keys sorted, empties last
  const int n_out = n_valid < k ? n_valid : k;

  if (tid < n_valid) {
    const double s = s_score[tid];
    const u32 r = key_row(s_out[tid]);
    int rank = 0;
    for (int j = 0; j < n_valid; ++j) {
      const double sj = s_score[j];
      const u32 rj = key_row(s_out[j]);
      rank += (sj > s) || (sj == s && rj < r);
    }
    if (rank < n_out) {
      out_scores[(size_t)qi * k + rank] = s;
      out_ids[(size_t)qi * k + rank] = (long long)r + row_offset;
    }
    if (rank == n_out - 1) {
      // certificate: every row outside the candidate set has fp32 score <= the fp32 score
      // of the last selected key, hence exact score <= that + eps.
      int flag = extra_flag;
      if (n_valid == KP) {
        const double last32 = (double)key_score(s_out[KP - 1]);
        if (!(s > last32 + cert_eps)) flag |= CMR_FLAG_UNCERTIFIED;
      }
      out_flags[qi] = flag;
    }
  }
  for (int i = n_out + tid; i < k; i += FIN_THREADS) {
    out_scores[(size_t)qi * k + i] = 0.0;
    out_ids[(size_t)qi * k + i] = -1;
  }
  if (tid == 0) {
    out_counts[qi] = n_out;
    if (n_out == 0) out_flags[qi] = extra_flag;
  }
}

// ---- dispatch between the two paths (dense.cu owns the C entry points) ----------------
struct DenseArgs {
  const uint16_t* emb;
  long long n_rows;
  int dim;
  const uint16_t* queries;
  int n_queries;
  int k;
  const uint8_t* row_mask;
  long long row_offset;
  double cert_eps;
  double* out_scores;
  long long* out_ids;
  int* out_counts;
  int* out_flags;
  void* workspace;
  size_t workspace_bytes;
  cudaStream_t stream;
};

// bm25.cu: select the k best (float64 score, id) keys over n_lists sorted lists per query
int launch_keyd_finalize(const KeyD* part, int n_lists, int n_queries, int kpl, long long row_offset, int k,
                         double* out_scores, long long* out_ids, int* out_counts, int* out_flags,
                         cudaStream_t st);

// dense_mma.cu
bool dense_mma_eligible(long long n_rows, int dim, int n_queries, int k, bool has_mask);
size_t dense_mma_workspace_bytes(long long n_rows, int dim, int n_queries, int k);
int dense_mma_topk(const DenseArgs& a);

}  // namespace cmr
