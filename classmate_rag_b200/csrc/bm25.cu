// bm25.cu -- A2: exact BM25 (Okapi) top-k over a CSR inverted index.
//
// Replaces BM25Okapi.get_scores + the full Python sort inside BM25Store.search
// (reference rag/retrieval/bm25.py:175-212; rank_bm25 semantics restated in
// oracle/np_oracle.py).  Two kernels:
//
//  bm25_tile_kernel      grid (query, tile group).  A CTA walks tiles of
//                        `tile_docs` consecutive documents and keeps their
//                        float64 accumulators in shared memory.  For every query
//                        token, in order, it streams the slice of that term's
//                        posting list that falls in the tile (located through the
//                        skip table, no search): coalesced reads of (doc, impact),
//                        acc[doc] += idf * impact.  Documents of one posting list
//                        are unique, so a term pass needs no atomics, and terms
//                        are separated by a barrier, so every document's additions
//                        happen in query-token order: the float64 result is
//                        bit-identical to rank_bm25's get_scores.  The accumulators
//                        never leave the SM: each warp scans a slice against its
//                        admission threshold and keeps a sorted list of the KP
//                        best (score, doc) keys.  ALL documents take part, so
//                        zero-score documents rank in ascending id order exactly
//                        like the reference's stable sort.  Every posting is read
//                        exactly once per query token.
//  bm25_finalize_kernel  one CTA per query: selects the k best keys over all CTA
//                        lists (score desc, doc asc) and writes them out.
//
// Algorithmic bytes per query: 12 * sum over query tokens of df(token)
// (int32 doc + float64 impact per posting).
#include "topk.cuh"

namespace cmr {

constexpr int BM_THREADS = 512;
constexpr int BM_WARPS = BM_THREADS / 32;
constexpr int BMF_THREADS = 1024;

__device__ __forceinline__ int ldg_stream_i32(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double ldg_stream_f64(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

template <int KPL>
__global__ void __launch_bounds__(BM_THREADS)
bm25_tile_kernel(cmr_lex_index ix, const int* __restrict__ q_terms, const int* __restrict__ q_ptr,
                 const uint8_t* __restrict__ row_mask, KeyD* __restrict__ part) {
  constexpr int KP = 32 * KPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);             // [tile_docs]
  KeyD* s_lists = reinterpret_cast<KeyD*>(acc + ix.tile_docs);   // [warps][KP]
  KeyD* s_out = s_lists + BM_WARPS * KP;                         // [KP]
  double* s_thr = reinterpret_cast<double*>(s_out + KP);         // [warps]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < BM_WARPS * KP; i += BM_THREADS) key_clear(s_lists[i]);
  if (tid < BM_WARPS) s_thr[tid] = -INFINITY;
  const int qlo = q_ptr[b], qhi = q_ptr[b + 1];
  KeyD* w_list = s_lists + warp * KP;
  volatile double* w_thr = s_thr + warp;

  for (int tile = blockIdx.y; tile < ix.n_tiles; tile += gridDim.y) {
    const long long tile_lo = (long long)tile * ix.tile_docs;
    const long long rem = ix.n_docs - tile_lo;
    const int n_here = rem < ix.tile_docs ? (int)rem : ix.tile_docs;
    for (int i = tid; i < ix.tile_docs; i += BM_THREADS) acc[i] = 0.0;
    __syncthreads();

    for (int j = qlo; j < qhi; ++j) {
      const int t = q_terms[j];
      if (t < 0 || t >= ix.n_terms) continue;  // unknown token contributes nothing (uniform)
      const double w = ix.idf[t];
      const long long base = ix.term_ptr[t];
      const uint32_t* sk = ix.tile_skip + (size_t)t * (ix.n_tiles + 1) + tile;
      const long long lo = base + sk[0], hi = base + sk[1];
      for (long long p = lo + tid; p < hi; p += BM_THREADS) {
        const int d = ldg_stream_i32(ix.post_doc + p) - (int)tile_lo;
        const double imp = ldg_stream_f64(ix.post_imp + p);
        acc[d] = __dadd_rn(acc[d], __dmul_rn(w, imp));  // no fma: rank_bm25 rounds the product
      }
      __syncthreads();
    }

    // per-warp selection over the tile's accumulators, ascending document order
    const int per_warp = ix.tile_docs / BM_WARPS;
    const int w_lo = warp * per_warp;
    int w_hi = w_lo + per_warp;
    if (w_hi > n_here) w_hi = n_here;
    for (int i0 = w_lo; i0 < w_hi; i0 += 32) {
      const int i = i0 + lane;
      double s = -INFINITY;
      bool ok = i < w_hi;
      if (ok) {
        s = acc[i];
        if (row_mask != nullptr) ok = row_mask[tile_lo + i] != 0;
      }
      // strict '>' is exact: documents arrive in ascending order, a tie with the
      // list's last entry loses the id tie-break
      unsigned bal = __ballot_sync(0xFFFFFFFFu, ok && (s > *w_thr));
      while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        KeyD key;
        key.s = __shfl_sync(0xFFFFFFFFu, s, src);
        key.id = (u32)(tile_lo + i0 + src);
        key.pad = 0;
        if (key.s > *w_thr) {
          KeyD new_last;
          key_clear(new_last);
          if (warp_list_insert<KP, KeyD>(w_list, key, lane, new_last) && !key_empty(new_last)) {
            if (lane == 0) *w_thr = new_last.s;
            __syncwarp();
          }
        }
      }
    }
    __syncthreads();  // accumulators are re-zeroed for the next tile
  }

  for (int i = tid; i < KP; i += BM_THREADS) key_clear(s_out[i]);
  __syncthreads();
  block_merge_lists<KP, KeyD>(s_lists, BM_WARPS, (size_t)KP, s_out, tid, BM_THREADS);
  __syncthreads();
  KeyD* dst = part + ((size_t)b * gridDim.y + blockIdx.y) * KP;
  for (int i = tid; i < KP; i += BM_THREADS) dst[i] = s_out[i];
}

template <int KPL>
__global__ void __launch_bounds__(BMF_THREADS)
bm25_finalize_kernel(const KeyD* __restrict__ part, int n_lists, long long row_offset, int k,
                     double* __restrict__ out_scores, long long* __restrict__ out_ids,
                     int* __restrict__ out_counts, int* __restrict__ out_flags) {
  constexpr int KP = 32 * KPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyD* s_heads = reinterpret_cast<KeyD*>(smem_raw);  // [n_lists + 1]
  constexpr int CAP = KP * KP < 4096 ? KP * KP : 4096;
  KeyD* s_buf = s_heads + n_lists + 1;                // [CAP]
  KeyD* s_out = s_buf + CAP;                          // [KP]
  int* s_cnt = reinterpret_cast<int*>(s_out + KP);
  const int qi = blockIdx.x, tid = threadIdx.x;
  block_select_from_lists<KP, CAP, KeyD>(part + (size_t)qi * n_lists * KP, n_lists, s_heads, s_buf, s_cnt,
                                    s_out, tid, BMF_THREADS);
  const int n_valid = count_valid(s_out, KP);
  const int n_out = n_valid < k ? n_valid : k;
  for (int i = tid; i < k; i += BMF_THREADS) {
    if (i < n_out) {
      out_scores[(size_t)qi * k + i] = s_out[i].s;
      out_ids[(size_t)qi * k + i] = (long long)s_out[i].id + row_offset;
    } else {
      out_scores[(size_t)qi * k + i] = 0.0;
      out_ids[(size_t)qi * k + i] = -1;
    }
  }
  if (tid == 0) {
    out_counts[qi] = n_out;
    out_flags[qi] = 0;
  }
}

struct Bm25Plan {
  int kpl;
  int grid_y;  // tile groups (lists per query)
  size_t smem_tile, smem_fin;
};

static int check_index(const cmr_lex_index* ix) {
  CMR_CHECK_ARG(ix != nullptr, "null index");
  CMR_CHECK_ARG(ix->n_docs >= 0 && ix->n_docs < 0xFFFFFFFFll, "n_docs out of range");
  CMR_CHECK_ARG(ix->tile_docs >= 512 && ix->tile_docs % 512 == 0, "tile_docs must be a positive multiple of 512");
  CMR_CHECK_ARG(ix->n_tiles >= 1 && (long long)ix->n_tiles * ix->tile_docs >= ix->n_docs, "n_tiles inconsistent with n_docs/tile_docs");
  CMR_CHECK_ARG(ix->n_terms >= 0, "n_terms negative");
  CMR_CHECK_ARG(ix->n_terms == 0 || (ix->term_ptr && ix->tile_skip && ix->idf), "null index arrays");
  return CMR_OK;
}

typedef void (*tile_fn_t)(cmr_lex_index, const int*, const int*, const uint8_t*, KeyD*);

static int make_plan(const cmr_lex_index& ix, int n_queries, int k, Bm25Plan* p) {
  p->kpl = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
  const int kp = 32 * p->kpl;
  p->smem_tile = (size_t)ix.tile_docs * 8 + (size_t)BM_WARPS * kp * 16 + (size_t)kp * 16 + BM_WARPS * 8 + 16;
  if (p->smem_tile > 220 * 1024) {
    set_error("bm25 tile shared memory %zu too large: lower tile_docs", p->smem_tile);
    return CMR_EUNSUPPORTED;
  }
  const int sms = sm_count();
  if (sms <= 0) return CMR_ECUDA;
  tile_fn_t fn = p->kpl == 1 ? bm25_tile_kernel<1> : (p->kpl == 2 ? bm25_tile_kernel<2> : bm25_tile_kernel<4>);
  struct Occ { tile_fn_t fn; size_t smem; int dev; int per_sm; };
  static Occ cache[32];
  static int n_cache = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 0;
  for (int i = 0; i < n_cache; ++i)
    if (cache[i].fn == fn && cache[i].smem == p->smem_tile && cache[i].dev == dev) per_sm = cache[i].per_sm;
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_tile);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(bm25_tile)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, BM_THREADS, p->smem_tile);
    if (e != cudaSuccess || per_sm <= 0) {
      set_error("occupancy query failed for bm25_tile (smem=%zu)", p->smem_tile);
      return CMR_ECUDA;
    }
    if (n_cache < 32) cache[n_cache++] = Occ{fn, p->smem_tile, dev, per_sm};
  }
  const long long resident = (long long)sms * per_sm;
  long long gy = (resident + n_queries - 1) / n_queries;
  if (gy < 1) gy = 1;
  if (gy > ix.n_tiles) gy = ix.n_tiles;
  if (gy > 65535) gy = 65535;
  p->grid_y = (int)gy;
  const int cap = kp * kp < 4096 ? kp * kp : 4096;
  p->smem_fin = (size_t)(p->grid_y + 1) * 16 + (size_t)cap * 16 + (size_t)kp * 16 + 16;
  if (p->smem_fin > 220 * 1024) {
    set_error("bm25 finalize shared memory %zu too large", p->smem_fin);
    return CMR_EUNSUPPORTED;
  }
  return CMR_OK;
}

template <int KPL>
static int launch_bm25(const cmr_lex_index& ix, const Bm25Plan& p, const int* q_terms, const int* q_ptr,
                       int n_queries, int k, const uint8_t* row_mask, long long row_offset,
                       double* out_scores, long long* out_ids, int* out_counts, int* out_flags,
                       KeyD* part, cudaStream_t st) {
  static int attr_dev_mask = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_dev_mask & (1 << dev))) {
    cudaError_t e = cudaFuncSetAttribute(bm25_finalize_kernel<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(bm25_finalize)");
    attr_dev_mask |= (1 << dev);
  }
  dim3 grid(n_queries, p.grid_y);
  bm25_tile_kernel<KPL><<<grid, BM_THREADS, p.smem_tile, st>>>(ix, q_terms, q_ptr, row_mask, part);
  bm25_finalize_kernel<KPL><<<n_queries, BMF_THREADS, p.smem_fin, st>>>(part, p.grid_y, row_offset, k, out_scores,
                                                                       out_ids, out_counts, out_flags);
  return CMR_OK;
}

}  // namespace cmr

using namespace cmr;

extern "C" size_t cmr_bm25_workspace_bytes(const cmr_lex_index* ix, int n_queries, int k) {
  if (check_index(ix) != CMR_OK || n_queries <= 0 || k <= 0 || k > CMR_MAX_K) return 0;
  Bm25Plan p;
  if (make_plan(*ix, n_queries, k, &p) != CMR_OK) return 0;
  return (size_t)n_queries * p.grid_y * (32 * p.kpl) * sizeof(KeyD);
}

extern "C" int cmr_bm25_topk(const cmr_lex_index* ix, const int32_t* q_terms, const int32_t* q_ptr,
                             int n_queries, int k, const uint8_t* row_mask, int64_t row_offset,
                             double* out_scores, int64_t* out_ids, int32_t* out_counts,
                             int32_t* out_flags, void* workspace, size_t workspace_bytes,
                             cmr_stream_t stream) {
  int rc = check_index(ix);
  if (rc != CMR_OK) return rc;
  CMR_CHECK_ARG(n_queries > 0 && n_queries <= 1 << 20, "n_queries %d out of range", n_queries);
  CMR_CHECK_ARG(k > 0 && k <= CMR_MAX_K, "k %d out of range (1..%d)", k, CMR_MAX_K);
  CMR_CHECK_ARG(q_ptr && out_scores && out_ids && out_counts && out_flags, "null pointer argument");
  Bm25Plan p;
  rc = make_plan(*ix, n_queries, k, &p);
  if (rc != CMR_OK) return rc;
  const size_t need = (size_t)n_queries * p.grid_y * (32 * p.kpl) * sizeof(KeyD);
  if (!workspace || workspace_bytes < need) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return CMR_EWORKSPACE;
  }
  CMR_CHECK_ARG(((uintptr_t)workspace % 16) == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  KeyD* part = (KeyD*)workspace;
  long long* ids = (long long*)out_ids;
  switch (p.kpl) {
    case 1: rc = launch_bm25<1>(*ix, p, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, ids, out_counts, out_flags, part, st); break;
    case 2: rc = launch_bm25<2>(*ix, p, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, ids, out_counts, out_flags, part, st); break;
    default: rc = launch_bm25<4>(*ix, p, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, ids, out_counts, out_flags, part, st); break;
  }
  if (rc != CMR_OK) return rc;
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}
