// dense.cu -- A1: exact dense top-k over the bf16 chunk-embedding matrix.
//
// Replaces the hnswlib search behind ChromaVectorStore.query
// (reference rag/retrieval/vector_chroma.py:204-253).  Two kernels:
//
//  dense_scan_kernel     HBM-streaming GEMV.  Every warp owns whole rows: a row
//                        of D bf16 is D/8 16-byte vectors, lane l loads vectors
//                        l, l+32, ... (fully coalesced 512 B per warp request,
//                        L1 no-allocate), multiplies with the query kept in
//                        registers as fp32, and a butterfly reduces the warp.
//                        The score never goes to memory: it is packed with the
//                        row into a 64-bit key and inserted into the warp's
//                        register-resident top-KP list only if it beats the
//                        list's current minimum (rare after warm-up).  At the
//                        end the CTA merges its warps' lists in shared memory
//                        and writes ONE sorted list of KP keys.
//  dense_finalize_kernel one CTA per query: selects the KP best keys over all
//                        CTA lists, rescoring each candidate exactly in float64
//                        (pinned order, bit-identical to the oracle), orders by
//                        (exact score desc, row asc) and writes the top k plus
//                        the over-selection certificate.
//
// Algorithmic bytes per query: n_rows * dim * 2 (the matrix is read once).
#include "topk.cuh"

namespace cmr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int FIN_THREADS = 1024;

// NV: 16-byte vectors per lane per row (dim <= NV*256); R: rows in flight per warp.
template <int NV, int R, int KPL>
__global__ void __launch_bounds__(SCAN_THREADS)
dense_scan_kernel(const uint4* __restrict__ emb, long long n_rows, int dim_vec,
                  const uint4* __restrict__ queries, const uint8_t* __restrict__ row_mask,
                  u64* __restrict__ part, long long rows_per_cta) {
  constexpr int KP = 32 * KPL;
  __shared__ u64 s_lists[SCAN_WARPS * KP];
  __shared__ u64 s_out[KP];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int qi = blockIdx.y;

  float qf[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = lane + 32 * j;
    uint4 qv = make_uint4(0, 0, 0, 0);
    if (v < dim_vec) qv = queries[(size_t)qi * dim_vec + v];
    qf[j][0] = bf16lo(qv.x); qf[j][1] = bf16hi(qv.x);
    qf[j][2] = bf16lo(qv.y); qf[j][3] = bf16hi(qv.y);
    qf[j][4] = bf16lo(qv.z); qf[j][5] = bf16hi(qv.z);
    qf[j][6] = bf16lo(qv.w); qf[j][7] = bf16hi(qv.w);
  }

  WarpList<KPL> list;
  list.init();

  const long long cta_lo = (long long)blockIdx.x * rows_per_cta;
  long long cta_hi = cta_lo + rows_per_cta;
  if (cta_hi > n_rows) cta_hi = n_rows;

  for (long long r0 = cta_lo + (long long)warp * R; r0 < cta_hi; r0 += (long long)SCAN_WARPS * R) {
    uint4 d[R][NV];
    bool valid[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const long long row = r0 + rr;
      valid[rr] = row < cta_hi;
      if (valid[rr] && row_mask != nullptr) valid[rr] = row_mask[row] != 0;
      const uint4* src = emb + (size_t)row * dim_vec;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int v = lane + 32 * j;
        d[rr][j] = make_uint4(0, 0, 0, 0);
        if (valid[rr] && v < dim_vec) d[rr][j] = ldg_stream(src + v);
      }
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const uint4 x = d[rr][j];
        acc0 = fmaf(bf16lo(x.x), qf[j][0], acc0);
        acc1 = fmaf(bf16hi(x.x), qf[j][1], acc1);
        acc0 = fmaf(bf16lo(x.y), qf[j][2], acc0);
        acc1 = fmaf(bf16hi(x.y), qf[j][3], acc1);
        acc0 = fmaf(bf16lo(x.z), qf[j][4], acc0);
        acc1 = fmaf(bf16hi(x.z), qf[j][5], acc1);
        acc0 = fmaf(bf16lo(x.w), qf[j][6], acc0);
        acc1 = fmaf(bf16hi(x.w), qf[j][7], acc1);
      }
      float s = acc0 + acc1;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
      if (valid[rr]) {  // warp-uniform
        const u64 key = make_key(s, (u32)(r0 + rr));
        if (key > list.kmin) list.insert(key, lane);
      }
    }
  }

  // CTA merge: warps' lists -> one sorted list of KP keys
  list.store(s_lists + warp * KP, lane);
  for (int i = threadIdx.x; i < KP; i += SCAN_THREADS) s_out[i] = 0ull;
  __syncthreads();
  block_merge_lists<KP>(s_lists, SCAN_WARPS, s_out, threadIdx.x, SCAN_THREADS);
  __syncthreads();
  u64* dst = part + ((size_t)qi * gridDim.x + blockIdx.x) * KP;
  for (int i = threadIdx.x; i < KP; i += SCAN_THREADS) dst[i] = s_out[i];
}

template <int KPL>
__global__ void __launch_bounds__(FIN_THREADS)
dense_finalize_kernel(const u64* __restrict__ part, int n_lists,
                      const uint16_t* __restrict__ emb, int dim,
                      const uint16_t* __restrict__ queries, long long row_offset, int k,
                      double cert_eps, double* __restrict__ out_scores,
                      long long* __restrict__ out_ids, int* __restrict__ out_counts,
                      int* __restrict__ out_flags) {
  constexpr int KP = 32 * KPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* s_heads = reinterpret_cast<u64*>(smem_raw);
  u64* s_stage = s_heads + n_lists;
  u64* s_out = s_stage + KP * KP;
  double* s_score = reinterpret_cast<double*>(s_out + KP);
  int* s_q = reinterpret_cast<int*>(s_score + KP);

  const int qi = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  block_select_from_lists<KP>(part + (size_t)qi * n_lists * KP, n_lists, s_heads, s_stage, s_q, s_out,
                              tid, FIN_THREADS);

  // exact rescoring, one warp per candidate
  const uint16_t* q = queries + (size_t)qi * dim;
  for (int c = warp; c < KP; c += FIN_THREADS / 32) {
    const u64 key = s_out[c];
    if (key == 0ull) continue;  // warp-uniform
    const uint16_t* row = emb + (size_t)key_row(key) * dim;
    const double s = warp_exact_dot(q, row, dim, lane);
    if (lane == 0) s_score[c] = s;
  }
  __syncthreads();

  // count valid (keys are sorted, zeros last)
  int n_valid = count_greater(s_out, KP, 0ull);
  const int n_out = n_valid < k ? n_valid : k;

  // final order: (exact score desc, row asc) by counting rank
  double kth_exact = 0.0;
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
      // certificate: every row outside the candidate set has fp32 score <= the
      // fp32 score of the last selected key, hence exact score <= that + eps.
      int flag = 0;
      if (n_valid == KP) {
        kth_exact = s;
        const double last32 = (double)key_score(s_out[KP - 1]);
        if (!(kth_exact > last32 + cert_eps)) flag = CMR_FLAG_UNCERTIFIED;
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
    if (n_out == 0) out_flags[qi] = 0;
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    __nv_bfloat16 b = __float2bfloat16_rn(src[i]);
    dst[i] = *reinterpret_cast<uint16_t*>(&b);
  }
}

// ---- host side -------------------------------------------------------------

struct DensePlan {
  int kpl;             // keys per lane (1, 2, 4)
  int nv;              // vectors per lane
  int grid_x;          // CTAs along rows
  long long rows_per_cta;
};

typedef void (*scan_fn_t)(const uint4*, long long, int, const uint4*, const uint8_t*, u64*, long long);

template <int KPL>
static scan_fn_t scan_fn_for_nv(int nv) {
  switch (nv) {
    case 1: return dense_scan_kernel<1, 4, KPL>;
    case 2: return dense_scan_kernel<2, 4, KPL>;
    case 3: return dense_scan_kernel<3, 4, KPL>;
    case 4: return dense_scan_kernel<4, 4, KPL>;
    case 5: case 6: return dense_scan_kernel<6, 2, KPL>;
    case 7: case 8: return dense_scan_kernel<8, 2, KPL>;
    default: return nullptr;
  }
}

static scan_fn_t scan_fn(int nv, int kpl) {
  return kpl == 1 ? scan_fn_for_nv<1>(nv) : (kpl == 2 ? scan_fn_for_nv<2>(nv) : scan_fn_for_nv<4>(nv));
}

// resident CTAs per SM of the scan kernel instance (persistent grid sizing)
static int scan_ctas_per_sm(int nv, int kpl) {
  static int cache[9][5] = {{0}};
  if (nv < 1 || nv > 8) return 0;
  if (cache[nv][kpl] == 0) {
    int n = 0;
    scan_fn_t fn = scan_fn(nv, kpl);
    if (!fn || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, SCAN_THREADS, 0) != cudaSuccess || n <= 0) {
      set_error("occupancy query failed for dense_scan<nv=%d,kpl=%d>", nv, kpl);
      return 0;
    }
    cache[nv][kpl] = n;
  }
  return cache[nv][kpl];
}

static int make_plan(long long n_rows, int dim, int k, DensePlan* p) {
  const int kp_needed = k + CMR_SLACK;
  p->kpl = kp_needed <= 32 ? 1 : (kp_needed <= 64 ? 2 : 4);
  const int dim_vec = dim / 8;
  p->nv = (dim_vec + 31) / 32;
  const int sms = sm_count();
  if (sms <= 0) return CMR_ECUDA;
  const int per_sm = scan_ctas_per_sm(p->nv, p->kpl);
  if (per_sm <= 0) return CMR_EUNSUPPORTED;
  long long ctas = (long long)sms * per_sm;
  // keep at least one full pass of work per warp group
  const long long min_rows = (long long)SCAN_WARPS * 4;
  long long max_ctas = (n_rows + min_rows - 1) / min_rows;
  if (max_ctas < 1) max_ctas = 1;
  if (ctas > max_ctas) ctas = max_ctas;
  p->grid_x = (int)ctas;
  long long rpc = (n_rows + ctas - 1) / ctas;
  // round rows_per_cta up to a whole number of warp-group steps for alignment
  p->rows_per_cta = rpc < 1 ? 1 : rpc;
  return CMR_OK;
}

static int launch_scan(const DensePlan& p, const uint16_t* emb, long long n_rows, int dim,
                       const uint16_t* queries, int n_queries, const uint8_t* row_mask, u64* part,
                       cudaStream_t st) {
  scan_fn_t fn = scan_fn(p.nv, p.kpl);
  if (!fn) {
    set_error("dim %d not supported (max 2048)", dim);
    return CMR_EUNSUPPORTED;
  }
  dim3 grid(p.grid_x, n_queries);
  fn<<<grid, SCAN_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(emb), n_rows, dim / 8,
                                    reinterpret_cast<const uint4*>(queries), row_mask, part,
                                    p.rows_per_cta);
  return CMR_OK;
}

template <int KPL>
static int launch_finalize(const DensePlan& p, const u64* part, const uint16_t* emb, int dim,
                           const uint16_t* queries, int n_queries, long long row_offset, int k,
                           double cert_eps, double* out_scores, long long* out_ids, int* out_counts,
                           int* out_flags, cudaStream_t st) {
  constexpr int KP = 32 * KPL;
  const size_t smem = (size_t)p.grid_x * 8 + (size_t)KP * KP * 8 + KP * 8 + KP * 8 + (KP + 1) * 4 + 16;
  static bool attr_set[5] = {false, false, false, false, false};
  if (!attr_set[KPL]) {
    cudaError_t e = cudaFuncSetAttribute(dense_finalize_kernel<KPL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(dense_finalize)");
    attr_set[KPL] = true;
  }
  if (smem > 200 * 1024) {
    set_error("finalize shared memory %zu too large", smem);
    return CMR_EUNSUPPORTED;
  }
  dense_finalize_kernel<KPL><<<n_queries, FIN_THREADS, smem, st>>>(
      part, p.grid_x, emb, dim, queries, row_offset, k, cert_eps, out_scores, out_ids, out_counts,
      out_flags);
  return CMR_OK;
}

}  // namespace cmr

using namespace cmr;

extern "C" size_t cmr_dense_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k) {
  DensePlan p;
  if (n_rows < 0 || dim <= 0 || n_queries <= 0 || k <= 0 || k > CMR_MAX_K) return 0;
  if (make_plan(n_rows, dim, k, &p) != CMR_OK) return 0;
  return (size_t)n_queries * p.grid_x * (32 * p.kpl) * sizeof(u64);
}

extern "C" int cmr_dense_topk(const uint16_t* emb, int64_t n_rows, int dim, const uint16_t* queries,
                              int n_queries, int k, const uint8_t* row_mask, int64_t row_offset,
                              double cert_eps, double* out_scores, int64_t* out_ids,
                              int32_t* out_counts, int32_t* out_flags, void* workspace,
                              size_t workspace_bytes, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_rows >= 0 && n_rows < 0xFFFFFFFFll, "n_rows %lld out of range", (long long)n_rows);
  CMR_CHECK_ARG(dim > 0 && dim % 8 == 0 && dim <= 2048, "dim %d must be a multiple of 8, <= 2048", dim);
  CMR_CHECK_ARG(n_queries > 0 && n_queries <= 65535, "n_queries %d out of range", n_queries);
  CMR_CHECK_ARG(k > 0 && k <= CMR_MAX_K, "k %d out of range (1..%d)", k, CMR_MAX_K);
  CMR_CHECK_ARG(queries && out_scores && out_ids && out_counts && out_flags, "null output/query pointer");
  CMR_CHECK_ARG(n_rows == 0 || emb, "null embedding matrix");
  CMR_CHECK_ARG(((uintptr_t)emb % 16) == 0 && ((uintptr_t)queries % 16) == 0, "emb/queries must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  DensePlan p;
  int rc = make_plan(n_rows, dim, k, &p);
  if (rc != CMR_OK) return rc;
  const size_t need = (size_t)n_queries * p.grid_x * (32 * p.kpl) * sizeof(u64);
  if (workspace_bytes < need || !workspace) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return CMR_EWORKSPACE;
  }
  u64* part = (u64*)workspace;
  const uint16_t* q = queries;
  long long* ids = (long long*)out_ids;
  switch (p.kpl) {
    case 1:
      rc = launch_scan(p, emb, n_rows, dim, q, n_queries, row_mask, part, st);
      if (rc == CMR_OK) rc = launch_finalize<1>(p, part, emb, dim, q, n_queries, row_offset, k, cert_eps, out_scores, ids, out_counts, out_flags, st);
      break;
    case 2:
      rc = launch_scan(p, emb, n_rows, dim, q, n_queries, row_mask, part, st);
      if (rc == CMR_OK) rc = launch_finalize<2>(p, part, emb, dim, q, n_queries, row_offset, k, cert_eps, out_scores, ids, out_counts, out_flags, st);
      break;
    default:
      rc = launch_scan(p, emb, n_rows, dim, q, n_queries, row_mask, part, st);
      if (rc == CMR_OK) rc = launch_finalize<4>(p, part, emb, dim, q, n_queries, row_offset, k, cert_eps, out_scores, ids, out_counts, out_flags, st);
      break;
  }
  if (rc != CMR_OK) return rc;
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_f32_to_bf16(const float* src, uint16_t* dst, int64_t n, cmr_stream_t stream) {
  CMR_CHECK_ARG(n >= 0 && (n == 0 || (src && dst)), "bad arguments");
  if (n == 0) return CMR_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}
