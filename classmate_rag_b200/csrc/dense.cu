// dense.cu -- A1: exact dense top-k over the bf16 chunk-embedding matrix.
//
// Replaces the hnswlib search behind ChromaVectorStore.query
// (reference rag/retrieval/vector_chroma.py:204-253).  Two kernels:
//
//  dense_scan_kernel     HBM-streaming scan, bound by HBM bandwidth.  A warp owns
//                        tiles of 16 consecutive rows.  Lane (g = lane/4,
//                        t4 = lane%4) issues 16-byte streaming loads of rows g and
//                        g+8 at column block t4 of every 32-column step, so one
//                        warp request covers 8 rows x 64 contiguous bytes (whole
//                        32 B sectors, every byte of the matrix fetched once).
//                        The loaded registers ARE the A fragments of
//                        mma.sync.m16n8k16 (the k index is permuted consistently
//                        on both operands, which a dot product does not see);
//                        the B fragments are the queries, staged once in shared
//                        memory.  The tensor pipe does the multiply-accumulate in
//                        fp32, so the SM issues ~9 instructions per row instead
//                        of ~130 for unpack+FFMA+shuffle, and one pass scores up
//                        to 8*NQ8 queries for the same bytes.  A ring of U units
//                        (U*1 KiB per warp) keeps loads in flight across tiles.
//                        Scores never go to memory: each is compared with the
//                        CTA's admission threshold and, rarely, inserted into the
//                        CTA-shared sorted list of KP keys in shared memory.
//  dense_finalize_kernel one CTA per query: selects the KP best keys over all
//                        CTA lists, rescoring each candidate exactly in float64
//                        (pinned order, bit-identical to the oracle), orders by
//                        (exact score desc, row asc) and writes the top k plus
//                        the over-selection certificate.
//
// Algorithmic bytes per scan pass: n_rows * dim * 2 (the matrix is read once,
// for up to 32 queries).
#include "dense_common.cuh"

namespace cmr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_U = 8;  // ring depth in 32-column units

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Rare path: at least one score of this tile reached one of the warp's thresholds.
template <int NQ8, int KP>
__device__ __forceinline__ void scan_tile_hits(const float (&c)[NQ8][4], long long r0, long long cta_hi,
                                               int q_base, int n_queries,
                                               const uint8_t* __restrict__ row_mask, u64* w_lists,
                                               float* w_thr, int lane) {
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int m = 0; m < NQ8; ++m) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int q = m * 8 + 2 * t4 + (r & 1);
      const long long row = r0 + g + ((r >> 1) ? 8 : 0);
      // strict '>' is exact: a warp sees its rows in ascending order, so a score that
      // only ties the list's last entry always loses the row tie-break
      bool hit = (c[m][r] > w_thr[q]) && (row < cta_hi) && (q_base + q < n_queries);
      if (hit && row_mask != nullptr) hit = row_mask[row] != 0;
      unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
      while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        const float s = __shfl_sync(0xFFFFFFFFu, c[m][r], src);
        const int qq = m * 8 + 2 * (src & 3) + (r & 1);
        const long long rr = r0 + (src >> 2) + ((r >> 1) ? 8 : 0);
        u64 new_last = 0ull;
        if (warp_list_insert<KP, u64>(w_lists + qq * KP, make_key(s, (u32)rr), lane, new_last) &&
            new_last != 0ull) {
          if (lane == 0) w_thr[qq] = key_score(new_last);
          __syncwarp();
        }
      }
    }
  }
}

// NQ8: query octets per pass; KPL: list entries per lane (KP = 32*KPL);
// COLS_FULL: dim is a multiple of 32*U columns, no column predicate needed.
template <int NQ8, int KPL, bool COLS_FULL>
__global__ void __launch_bounds__(SCAN_THREADS)
dense_scan_kernel(const uint4* __restrict__ emb, long long n_rows, int dim_vec, int steps_padded,
                  const uint4* __restrict__ queries, int n_queries,
                  const uint8_t* __restrict__ row_mask, u64* __restrict__ part,
                  long long rows_per_cta, int q_stride_vec) {
  constexpr int KP = 32 * KPL;
  constexpr int NQ = 8 * NQ8;
  constexpr int U = SCAN_U;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* s_q = reinterpret_cast<uint4*>(smem_raw);
  u64* s_lists = reinterpret_cast<u64*>(s_q + (size_t)NQ * q_stride_vec);  // [warp][query][KP]
  u64* s_out = s_lists + SCAN_WARPS * NQ * KP;                             // [KP]
  float* s_thr = reinterpret_cast<float*>(s_out + KP);                     // [warp][query]

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int q_base = blockIdx.y * NQ;

  // stage the queries of this pass (zero padded) and initialise the lists
  for (int idx = threadIdx.x; idx < NQ * q_stride_vec; idx += SCAN_THREADS) {
    const int qr = idx / q_stride_vec, v = idx - qr * q_stride_vec;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (q_base + qr < n_queries && v < dim_vec) val = queries[(size_t)(q_base + qr) * dim_vec + v];
    s_q[idx] = val;
  }
  for (int i = threadIdx.x; i < SCAN_WARPS * NQ * KP; i += SCAN_THREADS) s_lists[i] = 0ull;
  for (int i = threadIdx.x; i < SCAN_WARPS * NQ; i += SCAN_THREADS)
    s_thr[i] = (q_base + (i % NQ) < n_queries) ? -INFINITY : INFINITY;
  __syncthreads();
  u64* w_lists = s_lists + (size_t)warp * NQ * KP;
  float* w_thr = s_thr + warp * NQ;

  const long long cta_lo = (long long)blockIdx.x * rows_per_cta;
  long long cta_hi = cta_lo + rows_per_cta;
  if (cta_hi > n_rows) cta_hi = n_rows;

  long long r0 = cta_lo + (long long)warp * 16;
  if (r0 < cta_hi) {
    const long long last_row = cta_hi - 1;
    auto row_ptr = [&](long long r) {
      if (r > last_row) r = last_row;  // clamp: loads stay in bounds, result ignored
      return emb + (size_t)r * dim_vec + t4;
    };
    const uint4* pa = row_ptr(r0 + g);
    const uint4* pb = row_ptr(r0 + g + 8);
    uint4 xa[U], xb[U];
    auto load_unit = [&](const uint4* a, const uint4* b, int t, int slot) {
      if (COLS_FULL || (t * 4 + t4 < dim_vec)) {
        xa[slot] = ldg_stream(a + t * 4);
        xb[slot] = ldg_stream(b + t * 4);
      } else {
        xa[slot] = make_uint4(0, 0, 0, 0);
        xb[slot] = make_uint4(0, 0, 0, 0);
      }
    };
#pragma unroll
    for (int u = 0; u < U; ++u) load_unit(pa, pb, u, u);

    const uint4* sq_lane = s_q + (size_t)g * q_stride_vec + t4;

    while (r0 < cta_hi) {
      const long long r0n = r0 + (long long)SCAN_WARPS * 16;
      const bool has_next = r0n < cta_hi;
      const uint4* pan = row_ptr(r0n + g);
      const uint4* pbn = row_ptr(r0n + g + 8);
      float c[NQ8][4];
#pragma unroll
      for (int m = 0; m < NQ8; ++m) c[m][0] = c[m][1] = c[m][2] = c[m][3] = 0.f;

      for (int t0 = 0; t0 < steps_padded; t0 += U) {
        const bool last = (t0 + U >= steps_padded);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int t = t0 + u;
#pragma unroll
          for (int m = 0; m < NQ8; ++m) {
            const uint4 qv = sq_lane[(size_t)m * 8 * q_stride_vec + t * 4];
            mma_bf16_16816(c[m], xa[u].x, xb[u].x, xa[u].y, xb[u].y, qv.x, qv.y);
            mma_bf16_16816(c[m], xa[u].z, xb[u].z, xa[u].w, xb[u].w, qv.z, qv.w);
          }
          if (!last) load_unit(pa, pb, t + U, u);
          else if (has_next) load_unit(pan, pbn, u, u);
        }
      }

      // admission test against the warp's thresholds (fast path: 4*NQ8 compares)
      bool pass = false;
#pragma unroll
      for (int m = 0; m < NQ8; ++m) {
        const float2 th = *reinterpret_cast<const float2*>(w_thr + m * 8 + 2 * t4);
        pass |= (c[m][0] > th.x) | (c[m][2] > th.x) | (c[m][1] > th.y) | (c[m][3] > th.y);
      }
      if (__any_sync(0xFFFFFFFFu, pass))
        scan_tile_hits<NQ8, KP>(c, r0, cta_hi, q_base, n_queries, row_mask, w_lists, w_thr, lane);

      r0 = r0n;
      pa = pan;
      pb = pbn;
    }
  }

  // CTA merge: for every query, the warps' lists -> one sorted list of KP keys
  for (int q = 0; q < NQ; ++q) {
    __syncthreads();
    if (q_base + q >= n_queries) break;  // uniform
    for (int i = threadIdx.x; i < KP; i += SCAN_THREADS) s_out[i] = 0ull;
    __syncthreads();
    block_merge_lists<KP, u64>(s_lists + (size_t)q * KP, SCAN_WARPS, (size_t)NQ * KP, s_out, threadIdx.x,
                               SCAN_THREADS);
    __syncthreads();
    u64* dst = part + ((size_t)(q_base + q) * gridDim.x + blockIdx.x) * KP;
    for (int i = threadIdx.x; i < KP; i += SCAN_THREADS) dst[i] = s_out[i];
  }
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
  u64* s_heads = reinterpret_cast<u64*>(smem_raw);  // [n_lists + 1]
  constexpr int CAP = KP * KP < 4096 ? KP * KP : 4096;
  u64* s_buf = s_heads + n_lists + 1;               // [CAP]
  u64* s_out = s_buf + CAP;
  double* s_score = reinterpret_cast<double*>(s_out + KP);
  int* s_cnt = reinterpret_cast<int*>(s_score + KP);

  const int qi = blockIdx.x;
  const int tid = threadIdx.x;
  block_select_from_lists<KP, CAP, u64>(part + (size_t)qi * n_lists * KP, n_lists, s_heads, s_buf, s_cnt,
                                   s_out, tid, FIN_THREADS);

  dense_finalize_tail<KP>(s_out, s_score, emb, dim, queries + (size_t)qi * dim, row_offset, k, cert_eps, 0, qi,
                          out_scores, out_ids, out_counts, out_flags);
}

// ---- exhaustive exact scan (CMR_DENSE_EXACT) -----------------------------------------------
// Every row is scored with the exact float64 dot in the pinned order and ranked on that
// score, so there is nothing to certify: this is what a caller falls back to for the (rare)
// queries the fp32 passes flag as CMR_FLAG_UNCERTIFIED -- e.g. more exact duplicates of the
// k-th best row than the over-selection holds.  A warp owns rows w, w + W, ... in ascending
// order (strict '>' admission is then exact) and keeps a sorted list of the KP best
// (score, row) keys; about 40 instructions per lane and row: compute-bound (fp64 pipe), a few
// times slower than the streaming scan, never wrong.
constexpr int EXACT_THREADS = 256;
constexpr int EXACT_WARPS = EXACT_THREADS / 32;

template <int KPL>
__global__ void __launch_bounds__(EXACT_THREADS)
dense_exact_kernel(const uint16_t* __restrict__ emb, long long n_rows, int dim,
                   const uint16_t* __restrict__ queries, const uint8_t* __restrict__ row_mask,
                   KeyD* __restrict__ part) {
  constexpr int KP = 32 * KPL;
  __shared__ KeyD s_lists[EXACT_WARPS * KP];
  __shared__ KeyD s_out[KP];
  __shared__ double s_thr[EXACT_WARPS];
  const int qi = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < EXACT_WARPS * KP; i += EXACT_THREADS) key_clear(s_lists[i]);
  if (tid < EXACT_WARPS) s_thr[tid] = -INFINITY;
  __syncthreads();
  const uint16_t* q = queries + (size_t)qi * dim;
  KeyD* w_list = s_lists + warp * KP;
  volatile double* w_thr = s_thr + warp;
  const long long w0 = (long long)blockIdx.x * EXACT_WARPS + warp;
  const long long stride = (long long)gridDim.x * EXACT_WARPS;
  bool full = false;
  for (long long r = w0; r < n_rows; r += stride) {
    if (row_mask != nullptr && row_mask[r] == 0) continue;  // warp-uniform
    const double s = warp_exact_dot(q, emb + (size_t)r * dim, dim, lane);
    if (full && !(s > *w_thr)) continue;
    KeyD key;
    key.s = s;
    key.id = (u32)r;
    key.pad = 0;
    KeyD new_last;
    key_clear(new_last);
    if (warp_list_insert<KP, KeyD>(w_list, key, lane, new_last) && !key_empty(new_last)) {
      full = true;
      if (lane == 0) *w_thr = new_last.s;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < KP; i += EXACT_THREADS) key_clear(s_out[i]);
  __syncthreads();
  block_merge_lists<KP, KeyD>(s_lists, EXACT_WARPS, (size_t)KP, s_out, tid, EXACT_THREADS);
  __syncthreads();
  KeyD* dst = part + ((size_t)qi * gridDim.x + blockIdx.x) * KP;
  for (int i = tid; i < KP; i += EXACT_THREADS) dst[i] = s_out[i];
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

typedef void (*scan_fn_t)(const uint4*, long long, int, int, const uint4*, int, const uint8_t*, u64*,
                          long long, int);

struct DensePlan {
  int kpl;           // list entries per lane: KP = 32*kpl
  int nq8;           // query octets per pass
  int passes;        // grid.y
  int steps_padded;  // 32-column steps, rounded up to the ring depth
  int q_stride_vec;  // shared-memory stride of one staged query, in 16 B vectors
  bool cols_full;
  size_t smem;
  int grid_x;
  long long rows_per_cta;
  scan_fn_t fn;
};

template <int NQ8, int KPL>
static scan_fn_t pick_cols(bool full) {
  return full ? dense_scan_kernel<NQ8, KPL, true> : dense_scan_kernel<NQ8, KPL, false>;
}
template <int NQ8>
static scan_fn_t pick_kpl(int kpl, bool full) {
  return kpl == 1 ? pick_cols<NQ8, 1>(full) : (kpl == 2 ? pick_cols<NQ8, 2>(full) : pick_cols<NQ8, 4>(full));
}
static scan_fn_t pick_scan(int nq8, int kpl, bool full) {
  return nq8 == 1 ? pick_kpl<1>(kpl, full) : (nq8 == 2 ? pick_kpl<2>(kpl, full) : pick_kpl<4>(kpl, full));
}

static int make_plan(long long n_rows, int dim, int n_queries, int k, DensePlan* p) {
  const int kp_needed = k + CMR_SLACK;
  p->kpl = kp_needed <= 32 ? 1 : (kp_needed <= 64 ? 2 : 4);
  p->nq8 = n_queries <= 8 ? 1 : (n_queries <= 16 ? 2 : 4);
  if (p->kpl > 1) p->nq8 = 1;  // warp-private lists: keep them within shared memory
  p->passes = (n_queries + 8 * p->nq8 - 1) / (8 * p->nq8);
  const int steps = (dim + 31) / 32;
  p->steps_padded = (steps + SCAN_U - 1) / SCAN_U * SCAN_U;
  p->cols_full = (dim == p->steps_padded * 32);
  p->q_stride_vec = p->steps_padded * 4 + 4;  // +64 B: conflict-free LDS.128 across the 8 query rows
  const int nq = 8 * p->nq8, kp = 32 * p->kpl;
  p->smem = (size_t)nq * p->q_stride_vec * 16 + (size_t)SCAN_WARPS * nq * kp * 8 + (size_t)kp * 8 +
            (size_t)SCAN_WARPS * nq * 4;
  p->fn = pick_scan(p->nq8, p->kpl, p->cols_full);
  const int sms = sm_count();
  if (sms <= 0) return CMR_ECUDA;
  if (p->smem > 200 * 1024) {
    set_error("dense scan shared memory %zu too large", p->smem);
    return CMR_EUNSUPPORTED;
  }
  // occupancy of this (kernel, shared memory) pair, cached: the plan runs per call
  struct OccEntry { scan_fn_t fn; size_t smem; int dev; int per_sm; };
  static OccEntry occ_cache[64];
  static int occ_n = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 0;
  for (int i = 0; i < occ_n; ++i)
    if (occ_cache[i].fn == p->fn && occ_cache[i].smem == p->smem && occ_cache[i].dev == dev) per_sm = occ_cache[i].per_sm;
  if (per_sm == 0) {
    // opt in to the largest size any plan can ask for (the attribute is per kernel, not per
    // launch: setting it to this plan's size would lower it for an earlier, larger shape)
    cudaError_t e = cudaFuncSetAttribute(p->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(dense_scan)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p->fn, SCAN_THREADS, p->smem);
    if (e != cudaSuccess || per_sm <= 0) {
      set_error("occupancy query failed for dense_scan (nq8=%d kpl=%d smem=%zu)", p->nq8, p->kpl, p->smem);
      return CMR_ECUDA;
    }
    if (occ_n < 64) occ_cache[occ_n++] = OccEntry{p->fn, p->smem, dev, per_sm};
  }
  long long ctas = (long long)sms * per_sm;  // persistent: exactly the resident CTAs
  const long long min_rows = (long long)SCAN_WARPS * 16;
  long long max_ctas = (n_rows + min_rows - 1) / min_rows;
  if (max_ctas < 1) max_ctas = 1;
  if (ctas > max_ctas) ctas = max_ctas;
  p->grid_x = (int)ctas;
  long long rpc = (n_rows + ctas - 1) / ctas;
  rpc = (rpc + 15) / 16 * 16;  // whole 16-row tiles
  p->rows_per_cta = rpc < 16 ? 16 : rpc;
  return CMR_OK;
}

template <int KPL>
static int launch_finalize(const DensePlan& p, const u64* part, const uint16_t* emb, int dim,
                           const uint16_t* queries, int n_queries, long long row_offset, int k,
                           double cert_eps, double* out_scores, long long* out_ids, int* out_counts,
                           int* out_flags, cudaStream_t st) {
  constexpr int KP = 32 * KPL;
  constexpr int CAP = KP * KP < 4096 ? KP * KP : 4096;
  const size_t smem = (size_t)(p.grid_x + 1) * 8 + (size_t)CAP * 8 + KP * 8 + KP * 8 + 16;
  if (smem > 200 * 1024) {
    set_error("finalize shared memory %zu too large", smem);
    return CMR_EUNSUPPORTED;
  }
  static int attr_dev_mask = 0;  // per-device one-time opt-in to large dynamic shared memory
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_dev_mask & (1 << dev))) {
    cudaError_t e = cudaFuncSetAttribute(dense_finalize_kernel<KPL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(dense_finalize)");
    attr_dev_mask |= (1 << dev);
  }
  dense_finalize_kernel<KPL><<<n_queries, FIN_THREADS, smem, st>>>(
      part, p.grid_x, emb, dim, queries, row_offset, k, cert_eps, out_scores, out_ids, out_counts,
      out_flags);
  return CMR_OK;
}

}  // namespace cmr

using namespace cmr;

static size_t scan_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k) {
  DensePlan p;
  if (make_plan(n_rows, dim, n_queries, k, &p) != CMR_OK) return 0;
  return (size_t)n_queries * p.grid_x * (32 * p.kpl) * sizeof(u64);
}

static size_t exact_workspace_bytes(int64_t n_rows, int n_queries, int k);

extern "C" size_t cmr_dense_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k) {
  if (n_rows < 0 || dim <= 0 || dim % 8 != 0 || dim > 2048 || n_queries <= 0 || k <= 0 || k > CMR_MAX_K) {
    set_error("cmr_dense_workspace_bytes: bad shape");
    return 0;
  }
  // large enough for either path (the choice can depend on row_mask, known only at call time)
  size_t need = scan_workspace_bytes(n_rows, dim, n_queries, k);
  if (need == 0) return 0;
  if (dense_mma_eligible(n_rows, dim, n_queries, k, false)) {
    const size_t m = dense_mma_workspace_bytes(n_rows, dim, n_queries, k);
    if (m > need) need = m;
  }
  const size_t e = exact_workspace_bytes(n_rows, n_queries, k);
  if (e > need) need = e;
  return need;
}

static int dense_scan_topk(const DenseArgs& a) {
  cudaStream_t st = a.stream;
  DensePlan p;
  int rc = make_plan(a.n_rows, a.dim, a.n_queries, a.k, &p);
  if (rc != CMR_OK) return rc;
  const size_t need = (size_t)a.n_queries * p.grid_x * (32 * p.kpl) * sizeof(u64);
  if (a.workspace_bytes < need || !a.workspace) {
    set_error("workspace too small: %zu < %zu", a.workspace_bytes, need);
    return CMR_EWORKSPACE;
  }
  u64* part = (u64*)a.workspace;
  dim3 grid(p.grid_x, p.passes);
  p.fn<<<grid, SCAN_THREADS, p.smem, st>>>(reinterpret_cast<const uint4*>(a.emb), a.n_rows, a.dim / 8,
                                           p.steps_padded, reinterpret_cast<const uint4*>(a.queries),
                                           a.n_queries, a.row_mask, part, p.rows_per_cta, p.q_stride_vec);
  switch (p.kpl) {
    case 1: rc = launch_finalize<1>(p, part, a.emb, a.dim, a.queries, a.n_queries, a.row_offset, a.k, a.cert_eps, a.out_scores, a.out_ids, a.out_counts, a.out_flags, st); break;
    case 2: rc = launch_finalize<2>(p, part, a.emb, a.dim, a.queries, a.n_queries, a.row_offset, a.k, a.cert_eps, a.out_scores, a.out_ids, a.out_counts, a.out_flags, st); break;
    default: rc = launch_finalize<4>(p, part, a.emb, a.dim, a.queries, a.n_queries, a.row_offset, a.k, a.cert_eps, a.out_scores, a.out_ids, a.out_counts, a.out_flags, st); break;
  }
  if (rc != CMR_OK) return rc;
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

static int exact_grid(long long n_rows) {
  const int sms = sm_count();
  if (sms <= 0) return -1;
  long long ctas = (long long)sms * 8;
  const long long max_ctas = (n_rows + EXACT_WARPS - 1) / EXACT_WARPS;
  if (ctas > max_ctas) ctas = max_ctas;
  return (int)(ctas < 1 ? 1 : ctas);
}

static size_t exact_workspace_bytes(int64_t n_rows, int n_queries, int k) {
  const int kpl = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
  const int g = exact_grid(n_rows);
  return g < 0 ? 0 : (size_t)n_queries * g * (32 * kpl) * sizeof(KeyD);
}

static int dense_exact_topk(const DenseArgs& a) {
  const int kpl = a.k <= 32 ? 1 : (a.k <= 64 ? 2 : 4);
  const int g = exact_grid(a.n_rows);
  if (g < 0) return CMR_ECUDA;
  const size_t need = (size_t)a.n_queries * g * (32 * kpl) * sizeof(KeyD);
  if (a.workspace_bytes < need || !a.workspace) {
    set_error("workspace too small: %zu < %zu", a.workspace_bytes, need);
    return CMR_EWORKSPACE;
  }
  CMR_CHECK_ARG(((uintptr_t)a.workspace % 16) == 0, "workspace must be 16-byte aligned");
  CMR_CHECK_ARG(a.n_queries <= 65535, "exact scan takes at most 65535 queries per call");
  KeyD* part = (KeyD*)a.workspace;
  dim3 grid(g, a.n_queries);
  if (kpl == 1) dense_exact_kernel<1><<<grid, EXACT_THREADS, 0, a.stream>>>(a.emb, a.n_rows, a.dim, a.queries, a.row_mask, part);
  else if (kpl == 2) dense_exact_kernel<2><<<grid, EXACT_THREADS, 0, a.stream>>>(a.emb, a.n_rows, a.dim, a.queries, a.row_mask, part);
  else dense_exact_kernel<4><<<grid, EXACT_THREADS, 0, a.stream>>>(a.emb, a.n_rows, a.dim, a.queries, a.row_mask, part);
  int rc = launch_keyd_finalize(part, g, a.n_queries, kpl, a.row_offset, a.k, a.out_scores, a.out_ids, a.out_counts,
                                a.out_flags, a.stream);
  if (rc != CMR_OK) return rc;
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_dense_topk_ex(const uint16_t* emb, int64_t n_rows, int dim, const uint16_t* queries,
                                 int n_queries, int k, const uint8_t* row_mask, int64_t row_offset,
                                 double cert_eps, double* out_scores, int64_t* out_ids,
                                 int32_t* out_counts, int32_t* out_flags, void* workspace,
                                 size_t workspace_bytes, cmr_stream_t stream, int algo) {
  CMR_CHECK_ARG(n_rows >= 0 && n_rows < 0xFFFFFFFFll, "n_rows %lld out of range", (long long)n_rows);
  CMR_CHECK_ARG(dim > 0 && dim % 8 == 0 && dim <= 2048, "dim %d must be a multiple of 8, <= 2048", dim);
  CMR_CHECK_ARG(n_queries > 0 && n_queries <= 65535 * 8, "n_queries %d out of range", n_queries);
  CMR_CHECK_ARG(k > 0 && k <= CMR_MAX_K, "k %d out of range (1..%d)", k, CMR_MAX_K);
  CMR_CHECK_ARG(queries && out_scores && out_ids && out_counts && out_flags, "null output/query pointer");
  CMR_CHECK_ARG(n_rows == 0 || emb, "null embedding matrix");
  CMR_CHECK_ARG(((uintptr_t)emb % 16) == 0 && ((uintptr_t)queries % 16) == 0, "emb/queries must be 16-byte aligned");
  CMR_CHECK_ARG(algo == CMR_DENSE_AUTO || algo == CMR_DENSE_SCAN || algo == CMR_DENSE_MMA || algo == CMR_DENSE_EXACT,
                "unknown algo %d", algo);
  DenseArgs a{emb, (long long)n_rows, dim, queries, n_queries, k, row_mask, (long long)row_offset, cert_eps,
              out_scores, (long long*)out_ids, out_counts, out_flags, workspace, workspace_bytes,
              (cudaStream_t)stream};
  if (algo == CMR_DENSE_EXACT) return dense_exact_topk(a);
  const bool can_mma = dense_mma_eligible(n_rows, dim, n_queries, k, row_mask != nullptr);
  if (algo == CMR_DENSE_MMA && !can_mma) {
    set_error("tcgen05 path does not cover this shape (n_rows=%lld dim=%d B=%d k=%d mask=%d)",
              (long long)n_rows, dim, n_queries, k, row_mask != nullptr);
    return CMR_EUNSUPPORTED;
  }
  // auto: a single scan pass serves up to 8 queries at the HBM rate; above that the
  // tensor-core GEMM reads the matrix once for up to 128 queries (with or without a mask).
  const bool use_mma = algo == CMR_DENSE_MMA || (algo == CMR_DENSE_AUTO && can_mma && n_queries > 8);
  return use_mma ? dense_mma_topk(a) : dense_scan_topk(a);
}

extern "C" int cmr_dense_topk(const uint16_t* emb, int64_t n_rows, int dim, const uint16_t* queries,
                              int n_queries, int k, const uint8_t* row_mask, int64_t row_offset,
                              double cert_eps, double* out_scores, int64_t* out_ids,
                              int32_t* out_counts, int32_t* out_flags, void* workspace,
                              size_t workspace_bytes, cmr_stream_t stream) {
  return cmr_dense_topk_ex(emb, n_rows, dim, queries, n_queries, k, row_mask, row_offset, cert_eps,
                           out_scores, out_ids, out_counts, out_flags, workspace, workspace_bytes, stream,
                           CMR_DENSE_AUTO);
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
