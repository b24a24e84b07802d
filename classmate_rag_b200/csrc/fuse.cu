// fuse.cu -- the small latency-bound kernels that follow the two top-k scans:
//
//  gather_rows_kernel   copy candidate embedding rows into a dense buffer (rows
//                       of other shards are zero-filled, so an all-reduce SUM or
//                       an all-gather + merge over ranks reassembles the pool)
//  mmr_select_kernel    A4: greedy MMR re-ordering of the dense pool
//                       (reference rag/retrieval/fusion.py:39-61,80-102); the
//                       similarities are the pinned exact float64 dots
//  hybrid_fuse_kernel   A3+A5: Reciprocal Rank Fusion, per-id merge and the final
//                       stable sort of HybridRetriever.retrieve
//                       (rag/retrieval/fusion.py:17-36,108-167)
//  topk_merge_kernel    K7: merge the per-shard top-k lists gathered from the
//                       ranks (score desc, id asc)
//
// All float64 arithmetic uses explicit round-to-nearest intrinsics so that no
// fused multiply-add changes a rounding: results are bit-identical to Python.
#include "topk.cuh"

namespace cmr {

__global__ void gather_rows_kernel(const uint4* __restrict__ emb, long long n_rows, int dim_vec,
                                   long long row_offset, const long long* __restrict__ ids, int n_ids,
                                   uint4* __restrict__ out) {
  const int i = blockIdx.x;
  if (i >= n_ids) return;
  const long long local = ids[i] - row_offset;
  const bool mine = ids[i] >= 0 && local >= 0 && local < n_rows;
  for (int v = threadIdx.x; v < dim_vec; v += blockDim.x)
    out[(size_t)i * dim_vec + v] = mine ? emb[(size_t)local * dim_vec + v] : make_uint4(0, 0, 0, 0);
}

constexpr int MMR_THREADS = 512;
constexpr int MMR_MAX_POOL = 64;
constexpr int MMR_T = 4;   // a warp computes a 4 x 4 block of the similarity matrix at a time

// 4 x 4 block of exact float64 dots, rows (i0..i0+3) x (j0..j0+3), in warp_exact_dot's pinned order for every
// pair (lane l owns the 16-byte vectors l, l+32, ...; elements in order; halving tree).  float -> double
// conversion is the slow instruction here (16 lanes per clock and SM, the DFMA pipe does 64): a block
// converts every element once for 4 pairs instead of once per pair.  Rows beyond n_rows are clamped (their
// results are not used).  Lane 0 ends up with the 16 sums.
__device__ __forceinline__ void warp_exact_dot_block(const uint16_t* __restrict__ rows, int dim, int n_rows, int i0,
                                                     int j0, int lane, double (&acc)[MMR_T][MMR_T]) {
  const int nvec = dim >> 3;
  const uint4* ra[MMR_T];
  const uint4* rb[MMR_T];
#pragma unroll
  for (int t = 0; t < MMR_T; ++t) {
    const int i = i0 + t < n_rows ? i0 + t : n_rows - 1, j = j0 + t < n_rows ? j0 + t : n_rows - 1;
    ra[t] = reinterpret_cast<const uint4*>(rows + (size_t)i * dim);
    rb[t] = reinterpret_cast<const uint4*>(rows + (size_t)j * dim);
  }
#pragma unroll
  for (int r = 0; r < MMR_T; ++r)
#pragma unroll
    for (int c = 0; c < MMR_T; ++c) acc[r][c] = 0.0;
  for (int v = lane; v < nvec; v += 32) {
    u32 wa[MMR_T][4], wb[MMR_T][4];
#pragma unroll
    for (int t = 0; t < MMR_T; ++t) {
      const uint4 x = ra[t][v], y = rb[t][v];
      wa[t][0] = x.x, wa[t][1] = x.y, wa[t][2] = x.z, wa[t][3] = x.w;
      wb[t][0] = y.x, wb[t][1] = y.y, wb[t][2] = y.z, wb[t][3] = y.w;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double a[MMR_T], b[MMR_T];
#pragma unroll
        for (int t = 0; t < MMR_T; ++t) {
          a[t] = (double)(h ? bf16hi(wa[t][w]) : bf16lo(wa[t][w]));
          b[t] = (double)(h ? bf16hi(wb[t][w]) : bf16lo(wb[t][w]));
        }
#pragma unroll
        for (int r = 0; r < MMR_T; ++r)
#pragma unroll
          for (int c = 0; c < MMR_T; ++c) acc[r][c] = __fma_rn(a[r], b[c], acc[r][c]);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
    for (int r = 0; r < MMR_T; ++r)
#pragma unroll
      for (int c = 0; c < MMR_T; ++c) acc[r][c] = __dadd_rn(acc[r][c], __shfl_down_sync(0xFFFFFFFFu, acc[r][c], off));
}

// Index of the largest score among the lanes with `valid` (lowest index on equal scores), or -1: three
// warp-wide integer reductions on an order-preserving key instead of a shuffle tree of float64 compares.
__device__ __forceinline__ int warp_argbest(double score, int idx, bool valid) {
  const long long bits = __double_as_longlong(score + 0.0);   // + 0.0: -0.0 and 0.0 compare equal
  const unsigned long long key = (unsigned long long)(bits ^ ((bits >> 63) | (long long)0x8000000000000000ll));
  const u32 hi = (u32)(key >> 32), lo = (u32)key;
  const u32 mhi = __reduce_max_sync(0xFFFFFFFFu, valid ? hi : 0u);
  bool c = valid && hi == mhi;
  const u32 mlo = __reduce_max_sync(0xFFFFFFFFu, c ? lo : 0u);
  c = c && lo == mlo;
  const u32 mi = __reduce_min_sync(0xFFFFFFFFu, c ? (u32)idx : 0xFFFFFFFFu);
  return mi == 0xFFFFFFFFu ? -1 : (int)mi;
}

// One CTA per query.  cand_rows [B][pool][dim] bf16, cand_sims [B][pool] (exact
// q.c, i.e. the scores cmr_dense_topk returned), cand_ids [B][pool].
__global__ void __launch_bounds__(MMR_THREADS)
mmr_select_kernel(const uint16_t* __restrict__ cand_rows, const double* __restrict__ cand_sims,
                  const long long* __restrict__ cand_ids, const int* __restrict__ cand_counts, int pool,
                  int dim, int k, double lambda, long long* __restrict__ out_ids,
                  double* __restrict__ out_sims, int* __restrict__ out_counts, int stage_rows) {
  extern __shared__ __align__(16) unsigned char mmr_dyn[];   // the pool's rows when they fit (stage_rows)
  __shared__ double s_cc[MMR_MAX_POOL * MMR_MAX_POOL];
  __shared__ double s_q[MMR_MAX_POOL];
  __shared__ int s_sel[MMR_MAX_POOL];
  const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n = cand_counts[qi];
  if (n > pool) n = pool;
  const uint16_t* rows = cand_rows + (size_t)qi * pool * dim;
  for (int i = tid; i < n; i += MMR_THREADS) s_q[i] = cand_sims[(size_t)qi * pool + i];
  if (stage_rows) {
    // every row is read by n - 1 pairs: bring the pool in once, the dots then run at shared-memory latency
    const uint4* src = reinterpret_cast<const uint4*>(rows);
    uint4* dst = reinterpret_cast<uint4*>(mmr_dyn);
    for (int e = tid; e < n * (dim >> 3); e += MMR_THREADS) dst[e] = src[e];
    rows = reinterpret_cast<const uint16_t*>(mmr_dyn);
    __syncthreads();
  }
  // pairwise similarities (symmetric: the pinned order multiplies elementwise), one 4 x 4 block of the
  // upper triangle per warp and turn
  const int nb = (n + MMR_T - 1) / MMR_T, n_blocks = nb * (nb + 1) / 2;
  for (int p = warp; p < n_blocks; p += MMR_THREADS / 32) {
    int bi = 0, rem = p;
    while (rem >= nb - bi) {  // block row bi holds blocks (bi, bi..nb-1)
      rem -= nb - bi;
      ++bi;
    }
    const int bj = bi + rem;
    double acc[MMR_T][MMR_T];
    warp_exact_dot_block(rows, dim, n, bi * MMR_T, bj * MMR_T, lane, acc);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < MMR_T; ++r)
#pragma unroll
        for (int c = 0; c < MMR_T; ++c) {
          const int i = bi * MMR_T + r, j = bj * MMR_T + c;
          if (i < j && j < n) {
            s_cc[i * MMR_MAX_POOL + j] = acc[r][c];
            s_cc[j * MMR_MAX_POOL + i] = acc[r][c];
          }
        }
    }
  }
  __syncthreads();
  if (warp == 0) {
    // greedy selection by one warp: lane l owns candidates l and l+32 and keeps, for each, the running
    // maximum of its similarity to the items selected so far (one shared-memory read per pick); the best
    // (score desc, index asc -- the reference's strict '>' over ascending i) wins
    const int target = k < n ? k : n;
    int n_sel = 0;
    const int i0 = lane, i1 = lane + 32;
    bool live0 = i0 < n, live1 = i1 < n;
    const double q0 = live0 ? s_q[i0] : 0.0, q1 = live1 ? s_q[i1] : 0.0;
    double div0 = 0.0, div1 = 0.0;
    if (n > 0) {
      // np.argmax: lowest index on ties
      const bool one = live1 && (!live0 || q1 > q0);
      const int first = warp_argbest(one ? q1 : q0, one ? i1 : i0, live0 || live1);
      if (lane == 0) s_sel[0] = first;
      n_sel = 1;
      if (first == i0) live0 = false;
      if (first == i1) live1 = false;
      if (i0 < n) div0 = s_cc[i0 * MMR_MAX_POOL + first];
      if (i1 < n) div1 = s_cc[i1 * MMR_MAX_POOL + first];
    }
    const double one_minus = __dsub_rn(1.0, lambda);
    const double lq0 = __dmul_rn(lambda, q0), lq1 = __dmul_rn(lambda, q1);
    while (n_sel < target) {
      const double sc0 = __dsub_rn(lq0, __dmul_rn(one_minus, div0)), sc1 = __dsub_rn(lq1, __dmul_rn(one_minus, div1));
      const bool ok0 = live0 && sc0 > -1e9, ok1 = live1 && sc1 > -1e9;   // the reference starts its search at -1e9
      const bool one = ok1 && (!ok0 || sc1 > sc0);                        // the lower index keeps a tie
      const int best = warp_argbest(one ? sc1 : sc0, one ? i1 : i0, ok0 || ok1);
      if (best < 0) break;
      if (lane == 0) s_sel[n_sel] = best;
      ++n_sel;
      if (best == i0) live0 = false;
      if (best == i1) live1 = false;
      if (live0) {
        const double v = s_cc[i0 * MMR_MAX_POOL + best];
        if (v > div0) div0 = v;
      }
      if (live1) {
        const double v = s_cc[i1 * MMR_MAX_POOL + best];
        if (v > div1) div1 = v;
      }
    }
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
      if (i < n_sel) {
        out_ids[(size_t)qi * k + i] = cand_ids[(size_t)qi * pool + s_sel[i]];
        out_sims[(size_t)qi * k + i] = s_q[s_sel[i]];
      } else {
        out_ids[(size_t)qi * k + i] = -1;
        out_sims[(size_t)qi * k + i] = 0.0;
      }
    }
    if (lane == 0) out_counts[qi] = n_sel;
  }
}

constexpr int FUSE_THREADS = 128;
constexpr int FUSE_MAX_ITEMS = 256;

// One CTA per query.  vec list = dense results in their final (post-MMR) order.
__global__ void __launch_bounds__(FUSE_THREADS)
hybrid_fuse_kernel(const long long* __restrict__ vec_ids, const double* __restrict__ vec_sims,
                   const int* __restrict__ vec_counts, int kv, const long long* __restrict__ bm_ids,
                   const double* __restrict__ bm_scores, const int* __restrict__ bm_counts, int kb,
                   double w_vec, double w_bm, int rrf_k, int top_k, long long* __restrict__ out_ids,
                   double* __restrict__ out_fused, double* __restrict__ out_vdist,
                   double* __restrict__ out_bm25, int* __restrict__ out_counts) {
  __shared__ long long s_id[FUSE_MAX_ITEMS];
  __shared__ double s_fused[FUSE_MAX_ITEMS];
  __shared__ double s_vd[FUSE_MAX_ITEMS];   // vector_distance, NaN when absent
  __shared__ double s_bm[FUSE_MAX_ITEMS];   // bm25_score, NaN when absent
  __shared__ int s_n;
  const int qi = blockIdx.x, tid = threadIdx.x;
  int nv = vec_counts ? vec_counts[qi] : 0;
  if (nv > kv) nv = kv;
  int nb = (bm_counts && kb > 0) ? bm_counts[qi] : 0;
  if (nb > kb) nb = kb;
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  // vector items first, in order: fused = 0.0 + w*(1/(rrf_k+rank)) == the contribution
  for (int i = tid; i < nv; i += FUSE_THREADS) {
    s_id[i] = vec_ids[(size_t)qi * kv + i];
    s_fused[i] = __dmul_rn(w_vec, __ddiv_rn(1.0, (double)(rrf_k + i + 1)));
    s_vd[i] = __dsub_rn(1.0, vec_sims[(size_t)qi * kv + i]);
    s_bm[i] = nan;
  }
  if (tid == 0) s_n = nv;
  __syncthreads();
  // BM25 items: join on id (sequential: insertion order of BM25-only items matters)
  if (tid == 0) {
    int n = nv;
    for (int r = 0; r < nb; ++r) {
      const long long id = bm_ids[(size_t)qi * kb + r];
      const double contrib = __dmul_rn(w_bm, __ddiv_rn(1.0, (double)(rrf_k + r + 1)));
      int at = -1;
      for (int i = 0; i < n; ++i)
        if (s_id[i] == id) {
          at = i;
          break;
        }
      if (at < 0) {
        at = n++;
        s_id[at] = id;
        s_fused[at] = contrib;  // 0.0 + contrib
        s_vd[at] = nan;
      } else {
        s_fused[at] = __dadd_rn(s_fused[at], contrib);
      }
      s_bm[at] = bm_scores[(size_t)qi * kb + r];
    }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  const int n_out = n < top_k ? n : top_k;
  // stable descending sort on (fused, -vector_distance); absent distance -> -0.0
  for (int i = tid; i < n; i += FUSE_THREADS) {
    const double f = s_fused[i];
    const double t = (s_vd[i] != s_vd[i]) ? 0.0 : -s_vd[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double fj = s_fused[j];
      const double tj = (s_vd[j] != s_vd[j]) ? 0.0 : -s_vd[j];
      const bool greater = (fj > f) || (fj == f && tj > t);
      const bool equal = (fj == f) && (tj == t);
      rank += greater || (equal && j < i);
    }
    if (rank < n_out) {
      out_ids[(size_t)qi * top_k + rank] = s_id[i];
      out_fused[(size_t)qi * top_k + rank] = f;
      out_vdist[(size_t)qi * top_k + rank] = s_vd[i];
      out_bm25[(size_t)qi * top_k + rank] = s_bm[i];
    }
  }
  for (int i = n_out + tid; i < top_k; i += FUSE_THREADS) {
    out_ids[(size_t)qi * top_k + i] = -1;
    out_fused[(size_t)qi * top_k + i] = 0.0;
    out_vdist[(size_t)qi * top_k + i] = nan;
    out_bm25[(size_t)qi * top_k + i] = nan;
  }
  if (tid == 0) out_counts[qi] = n_out;
}

constexpr int MERGE_THREADS = 256;

// in_* are laid out [G][B][k]; one CTA per query.
__global__ void __launch_bounds__(MERGE_THREADS)
topk_merge_kernel(const double* __restrict__ in_scores, const long long* __restrict__ in_ids,
                  const int* __restrict__ in_counts, int n_parts, int n_queries, int k,
                  double* __restrict__ out_scores, long long* __restrict__ out_ids,
                  int* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_s = reinterpret_cast<double*>(smem_raw);          // [G*k]
  long long* s_i = reinterpret_cast<long long*>(s_s + n_parts * k);
  __shared__ int s_total;
  const int qi = blockIdx.x, tid = threadIdx.x;
  const int n = n_parts * k;
  if (tid == 0) s_total = 0;
  __syncthreads();
  for (int e = tid; e < n; e += MERGE_THREADS) {
    const int g = e / k, r = e - g * k;
    const int cnt = in_counts[(size_t)g * n_queries + qi];
    const size_t src = ((size_t)g * n_queries + qi) * k + r;
    const bool valid = r < cnt && in_ids[src] >= 0;
    s_s[e] = valid ? in_scores[src] : 0.0;
    s_i[e] = valid ? in_ids[src] : -1;
    if (valid) atomicAdd(&s_total, 1);
  }
  __syncthreads();
  const int n_out = s_total < k ? s_total : k;
  for (int e = tid; e < n; e += MERGE_THREADS) {
    const long long id = s_i[e];
    if (id < 0) continue;
    const double s = s_s[e];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const long long idj = s_i[j];
      if (idj < 0) continue;
      const double sj = s_s[j];
      rank += (sj > s) || (sj == s && idj < id);
    }
    if (rank < n_out) {
      out_scores[(size_t)qi * k + rank] = s;
      out_ids[(size_t)qi * k + rank] = id;
    }
  }
  for (int i = n_out + tid; i < k; i += MERGE_THREADS) {
    out_scores[(size_t)qi * k + i] = 0.0;
    out_ids[(size_t)qi * k + i] = -1;
  }
  if (tid == 0) out_counts[qi] = n_out;
}

// ---- K7, single exchange ------------------------------------------------------------------
// On a row-sharded corpus a step needs the other shards' dense pool (scores, ids AND the
// embedding rows the MMR step reads) and their BM25 lists.  shard_pack_kernel writes all of it
// into one message per rank, so the step has ONE collective (an all-gather of the messages);
// shard_merge_kernel then merges the G messages per query.  Message of one query (bytes):
//   [dense scores f64 x pool][dense ids i64 x pool][bm scores f64 x kb][bm ids i64 x kb]
//   [int32 x 4: dense count, dense flag, bm count, 0][rows bf16 x pool x dim]
__host__ __device__ inline size_t shard_msg_bytes(int pool, int kb, int dim) {
  return (size_t)16 * pool + (size_t)16 * kb + 16 + (size_t)2 * pool * dim;
}

constexpr int SHARD_THREADS = 256;

// Peer-memory form of the exchange (ShardP2P below): instead of one local message that NCCL
// then all-gathers, the kernel stores every element straight into the receive buffer of EVERY
// rank (its own included) over NVLink -- peer pointers from a symmetric-memory rendezvous --
// and the last CTA to finish publishes an epoch flag on every rank (release, system scope).
// shard_merge_kernel on each rank waits for all flags of the epoch (acquire) and merges.  No
// collective launch; the transfer overlaps the packing query by query.
struct ShardP2P {
  const unsigned long long* peer_recv;   // device array [n_parts]: base of every rank's receive buffer
  const unsigned long long* peer_flags;  // device array [n_parts]: base of every rank's flag words (u32)
  unsigned int* state;                   // local: [0] epoch of the last completed pack, [1] CTAs done
  int n_parts, my_rank;
  unsigned long long slot_stride;        // bytes between the slots of two source ranks
  unsigned long long parity_stride;      // bytes between the two buffers used on alternating steps
  const unsigned long long* peer_rows;   // pull form: device array [n_parts] of every rank's row matrix, or null
  const long long* peer_row_lo;          // pull form: device array [n_parts], global id of every rank's first row
};

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool P2P>
__global__ void __launch_bounds__(SHARD_THREADS)
shard_pack_kernel(const double* __restrict__ d_scores, const long long* __restrict__ d_ids,
                  const int* __restrict__ d_counts, const int* __restrict__ d_flags, int pool,
                  const double* __restrict__ b_scores, const long long* __restrict__ b_ids,
                  const int* __restrict__ b_counts, int kb, const uint4* __restrict__ emb, long long n_rows,
                  int dim, long long row_offset, unsigned char* __restrict__ msg, ShardP2P x) {
  const int qi = blockIdx.x, tid = threadIdx.x;
  const size_t mb = shard_msg_bytes(pool, kb, dim);
  const size_t o_ids = (size_t)8 * pool, o_bs = (size_t)16 * pool, o_bi = o_bs + (size_t)8 * kb,
               o_hdr = o_bs + (size_t)16 * kb, o_rows = o_hdr + 16;
  const int n_dst = P2P ? x.n_parts : 1;
  unsigned int epoch = 0;
  if (P2P) epoch = x.state[0] + 1u;
  // destination g of this query's message
  auto dst = [&](int g) -> unsigned char* {
    if (!P2P) return msg + (size_t)qi * mb;
    return reinterpret_cast<unsigned char*>(x.peer_recv[g]) + (size_t)(epoch & 1u) * x.parity_stride +
           (size_t)x.my_rank * x.slot_stride + (size_t)qi * mb;
  };
  for (int i = tid; i < pool; i += SHARD_THREADS) {
    const double sv = d_scores[(size_t)qi * pool + i];
    const long long iv = d_ids[(size_t)qi * pool + i];
    for (int g = 0; g < n_dst; ++g) {
      unsigned char* m = dst(g);
      reinterpret_cast<double*>(m)[i] = sv;
      reinterpret_cast<long long*>(m + o_ids)[i] = iv;
    }
  }
  for (int i = tid; i < kb; i += SHARD_THREADS) {
    const double sv = b_scores[(size_t)qi * kb + i];
    const long long iv = b_ids[(size_t)qi * kb + i];
    for (int g = 0; g < n_dst; ++g) {
      unsigned char* m = dst(g);
      reinterpret_cast<double*>(m + o_bs)[i] = sv;
      reinterpret_cast<long long*>(m + o_bi)[i] = iv;
    }
  }
  if (tid == 0) {
    const int4 hdr = make_int4(d_counts[qi], d_flags != nullptr ? d_flags[qi] : 0, kb > 0 ? b_counts[qi] : 0, 0);
    for (int g = 0; g < n_dst; ++g) *reinterpret_cast<int4*>(dst(g) + o_hdr) = hdr;
  }
  if (dim > 0) {
    const int dim_vec = dim / 8;
    const int cnt = d_counts[qi];
    for (int e = tid; e < pool * dim_vec; e += SHARD_THREADS) {
      const int r = e / dim_vec, v = e - r * dim_vec;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r < cnt) {
        const long long local = d_ids[(size_t)qi * pool + r] - row_offset;
        if (local >= 0 && local < n_rows) val = emb[(size_t)local * dim_vec + v];
      }
      for (int g = 0; g < n_dst; ++g) reinterpret_cast<uint4*>(dst(g) + o_rows)[e] = val;
    }
  }
  if (P2P) {
    // publish: every thread's peer stores are ordered before the CTA's arrival; the last CTA
    // raises the epoch flag of this (source rank, parity) on every rank
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned int arrived = atomicAdd(&x.state[1], 1u);
      if (arrived == gridDim.x - 1) {
        __threadfence_system();
        for (int g = 0; g < x.n_parts; ++g) {
          unsigned int* f = reinterpret_cast<unsigned int*>(x.peer_flags[g]) + (epoch & 1u) * x.n_parts + x.my_rank;
          st_release_sys_u32(f, epoch);
        }
        x.state[1] = 0u;
        __threadfence();
        x.state[0] = epoch;   // the merge kernel (next in the stream) reads the epoch from here
      }
    }
  }
}

// One CTA per query over the gathered messages [G][B][msg].  Ranking by counting over the
// G*pool (G*kb) entries with the total order (score desc, id asc): identical for any G.
__global__ void __launch_bounds__(SHARD_THREADS)
shard_merge_kernel(const unsigned char* gathered /* peers store into it until the flags are up: no __restrict__,
                   so its loads are ordinary ld.global after the acquire, never ld.global.nc */, size_t slot_stride,
                   const unsigned int* __restrict__ wait_flags, const unsigned int* __restrict__ wait_state,
                   unsigned long long parity_stride, int* __restrict__ timeout_flag,
                   int n_parts, int n_queries, int pool, int kb,
                   int dim, int msg_dim /* row width the messages were packed with: dim, or 0 when rows are pulled */,
                   const unsigned long long* __restrict__ peer_rows, const long long* __restrict__ peer_row_lo,
                   double* __restrict__ d_scores, long long* __restrict__ d_ids,
                   int* __restrict__ d_counts, int* __restrict__ d_flags, uint4* __restrict__ d_rows,
                   double* __restrict__ b_scores, long long* __restrict__ b_ids, int* __restrict__ b_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_d = n_parts * pool, n_b = n_parts * kb;
  double* s_s = reinterpret_cast<double*>(smem_raw);        // [max(n_d, n_b)]
  long long* s_i = reinterpret_cast<long long*>(s_s + (n_d > n_b ? n_d : n_b));
  int* s_src = reinterpret_cast<int*>(s_i + (n_d > n_b ? n_d : n_b));  // [pool] entry feeding each output slot
  __shared__ int s_total, s_flag;
  const int qi = blockIdx.x, tid = threadIdx.x;
  const size_t mb = shard_msg_bytes(pool, kb, msg_dim);
  if (wait_flags != nullptr) {
    // peer-memory exchange: wait until every rank has published this step's message
    const unsigned int epoch = wait_state[0];
    gathered += (size_t)(epoch & 1u) * parity_stride;
    if (tid < n_parts) {
      const unsigned int* f = wait_flags + (epoch & 1u) * n_parts + tid;
      const long long t0 = clock64();
      while (ld_acquire_sys_u32(f) != epoch) {
        if (clock64() - t0 > 20000000000ll) {  // ~10 s: a rank died; fail loudly instead of hanging
          *timeout_flag = 1;
          break;
        }
      }
    }
    __syncthreads();
  }
  auto msg_of = [&](int g) { return gathered + (size_t)g * slot_stride + (size_t)qi * mb; };

  for (int pass = 0; pass < 2; ++pass) {
    const int k = pass == 0 ? pool : kb;
    if (k == 0) continue;
    const int n = n_parts * k;
    if (tid == 0) {
      s_total = 0;
      s_flag = 0;
    }
    for (int i = tid; i < pool; i += SHARD_THREADS) s_src[i] = -1;
    __syncthreads();
    for (int e = tid; e < n; e += SHARD_THREADS) {
      const int g = e / k, r = e - g * k;
      const unsigned char* m = msg_of(g);
      const int* hdr = reinterpret_cast<const int*>(m + (size_t)16 * pool + (size_t)16 * kb);
      const double* sc = reinterpret_cast<const double*>(m + (pass == 0 ? 0 : (size_t)16 * pool));
      const long long* id = reinterpret_cast<const long long*>(m + (pass == 0 ? (size_t)8 * pool
                                                                                : (size_t)16 * pool + (size_t)8 * kb));
      const int cnt = hdr[pass == 0 ? 0 : 2];
      const bool valid = r < cnt && id[r] >= 0;
      s_s[e] = valid ? sc[r] : 0.0;
      s_i[e] = valid ? id[r] : -1;
      if (valid) atomicAdd(&s_total, 1);
      if (pass == 0 && r == 0 && hdr[1]) atomicOr(&s_flag, hdr[1]);
    }
    __syncthreads();
    const int n_out = s_total < k ? s_total : k;
    double* o_s = pass == 0 ? d_scores : b_scores;
    long long* o_i = pass == 0 ? d_ids : b_ids;
    for (int e = tid; e < n; e += SHARD_THREADS) {
      const long long id = s_i[e];
      if (id < 0) continue;
      const double sv = s_s[e];
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const long long idj = s_i[j];
        if (idj < 0) continue;
        const double sj = s_s[j];
        rank += (sj > sv) || (sj == sv && idj < id);
      }
      if (rank < n_out) {
        o_s[(size_t)qi * k + rank] = sv;
        o_i[(size_t)qi * k + rank] = id;
        if (pass == 0) s_src[rank] = e;
      }
    }
    for (int i = n_out + tid; i < k; i += SHARD_THREADS) {
      o_s[(size_t)qi * k + i] = 0.0;
      o_i[(size_t)qi * k + i] = -1;
    }
    if (tid == 0) {
      (pass == 0 ? d_counts : b_counts)[qi] = n_out;
      if (pass == 0) d_flags[qi] = s_flag;
    }
    __syncthreads();
    if (pass == 0 && dim > 0 && d_rows != nullptr) {
      const int dim_vec = dim / 8, total = pool * dim_vec;
      if (peer_rows != nullptr) {
        // pull: every row of the merged pool lives in its source rank's matrix (mapped here).  The loads
        // cross NVLink (microseconds each), so a thread keeps PULL_U of them in flight before storing.
        constexpr int PULL_U = 12;
        for (int e0 = tid; e0 < total; e0 += SHARD_THREADS * PULL_U) {
          uint4 val[PULL_U];
#pragma unroll
          for (int u = 0; u < PULL_U; ++u) {
            const int e = e0 + u * SHARD_THREADS;
            val[u] = make_uint4(0, 0, 0, 0);
            if (e < total) {
              const int r = e / dim_vec, v = e - r * dim_vec;
              const int src = s_src[r];
              if (src >= 0) {
                const int g = src / pool;
                const long long local = s_i[src] - peer_row_lo[g];
                val[u] = reinterpret_cast<const uint4*>(peer_rows[g])[(size_t)local * dim_vec + v];
              }
            }
          }
#pragma unroll
          for (int u = 0; u < PULL_U; ++u) {
            const int e = e0 + u * SHARD_THREADS;
            if (e < total) d_rows[(size_t)qi * total + e] = val[u];
          }
        }
      } else {
        for (int e = tid; e < total; e += SHARD_THREADS) {
          const int r = e / dim_vec, v = e - r * dim_vec;
          uint4 val = make_uint4(0, 0, 0, 0);
          const int src = s_src[r];
          if (src >= 0) {
            const int g = src / pool, rr = src - g * pool;
            const uint4* rows = reinterpret_cast<const uint4*>(msg_of(g) + (size_t)16 * pool + (size_t)16 * kb + 16);
            val = rows[(size_t)rr * dim_vec + v];
          }
          d_rows[(size_t)qi * total + e] = val;
        }
      }
      __syncthreads();
    }
  }
}

// A3 standalone: rrf_fuse over any number of rank lists (rag/retrieval/fusion.py:17-36).
// One CTA.  Thread i owns entry i of the concatenated lists; an entry that is the first
// occurrence of its id walks every list in order and adds w * (1.0 / (rrf_k + rank)) exactly
// as the reference's dict accumulation does (list by list, rank by rank).  Output keeps the
// dict's insertion order (first appearance).
constexpr int RRF_THREADS = 256;
constexpr int RRF_MAX_ENTRIES = 4096;

__global__ void __launch_bounds__(RRF_THREADS)
rrf_fuse_kernel(const long long* __restrict__ ids, const int* __restrict__ counts, int n_lists, int max_len,
                const double* __restrict__ weights, int rrf_k, long long* __restrict__ out_ids,
                double* __restrict__ out_scores, int* __restrict__ out_count) {
  __shared__ unsigned char s_first[RRF_MAX_ENTRIES];
  __shared__ int s_pos[RRF_MAX_ENTRIES];
  const int tid = threadIdx.x;
  const int total = n_lists * max_len;
  for (int e = tid; e < total; e += RRF_THREADS) {
    const int l = e / max_len, r = e - l * max_len;
    bool first = r < counts[l];
    if (first) {
      const long long id = ids[e];
      for (int l2 = 0; l2 <= l && first; ++l2) {
        const int lim = l2 < l ? counts[l2] : r;
        for (int r2 = 0; r2 < lim; ++r2)
          if (ids[(size_t)l2 * max_len + r2] == id) {
            first = false;
            break;
          }
      }
    }
    s_first[e] = first ? 1 : 0;
  }
  __syncthreads();
  if (tid == 0) {  // exclusive scan over <= 4096 flags: output slot of every first occurrence
    int n = 0;
    for (int e = 0; e < total; ++e) {
      s_pos[e] = n;
      n += s_first[e];
    }
    *out_count = n;
  }
  __syncthreads();
  for (int e = tid; e < total; e += RRF_THREADS) {
    if (!s_first[e]) continue;
    const long long id = ids[e];
    double score = 0.0;
    for (int l = 0; l < n_lists; ++l) {
      const double w = weights[l];
      for (int r = 0; r < counts[l]; ++r)
        if (ids[(size_t)l * max_len + r] == id)
          score = __dadd_rn(score, __dmul_rn(w, __ddiv_rn(1.0, (double)(rrf_k + (r + 1)))));
    }
    out_ids[s_pos[e]] = id;
    out_scores[s_pos[e]] = score;
  }
}

// N1: metadata `where` evaluated on the device.  Every filterable field is a dictionary-coded
// int32 column ([n_fields][n_rows], -1 = field absent); a clause list is a conjunction of
// (field, code) equalities.  code -1 asks for "field absent" (the reference's BM25 filter
// matches a None-valued filter key only against documents lacking the field, quirk Q1);
// the host encodes a value that is not in the dictionary as code -2, which matches nothing.
// `alive` (optional) carries tombstones of deleted rows.
__global__ void filter_mask_kernel(const int* __restrict__ cols, long long n_rows, const int* __restrict__ clause_field,
                                   const int* __restrict__ clause_code, int n_clauses,
                                   const uint8_t* __restrict__ alive, uint8_t* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n_rows; i += stride) {
    bool ok = alive == nullptr || alive[i] != 0;
    for (int c = 0; c < n_clauses && ok; ++c)
      ok = cols[(size_t)clause_field[c] * n_rows + i] == clause_code[c];
    out[i] = ok ? 1 : 0;
  }
}

// N1, statistics of a filtered BM25 search (rag/retrieval/bm25.py:184-191 rebuilds BM25Okapi over the
// filtered entries, so df / N / avgdl / idf are the SUBSET's).  One pass over the CSR: a warp takes
// MDF_CHUNK consecutive postings, finds the term its first posting belongs to (bisection on term_ptr) and
// walks the term boundaries inside the chunk; per (term, chunk) piece the lanes sum mask[post_doc[i]] with
// coalesced loads and one lane adds the count to df[t] and lowers first[t] to the first passing posting.
// Bytes: 4 per posting + the mask gathers (L2-resident: one byte per document).
constexpr int MDF_THREADS = 256;
constexpr int MDF_CHUNK = 4096;

__global__ void __launch_bounds__(MDF_THREADS)
masked_df_kernel(const long long* __restrict__ term_ptr, const int* __restrict__ post_doc, int n_terms,
                 long long n_post, const uint8_t* __restrict__ mask, int* __restrict__ df, int* __restrict__ first) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * MDF_THREADS + threadIdx.x) >> 5;
  const long long lo = warp * MDF_CHUNK;
  if (lo >= n_post) return;
  const long long hi = lo + MDF_CHUNK < n_post ? lo + MDF_CHUNK : n_post;
  // largest t with term_ptr[t] <= lo (empty terms before it are skipped by the walk below)
  int a = 0, b = n_terms;
  while (b - a > 1) {
    const int mid = (a + b) >> 1;
    if (term_ptr[mid] <= lo) a = mid;
    else b = mid;
  }
  int t = a;
  long long pos = lo;
  while (pos < hi && t < n_terms) {
    const long long t_end = term_ptr[t + 1];
    if (t_end <= pos) {
      ++t;
      continue;
    }
    const long long e = t_end < hi ? t_end : hi;
    int sum = 0;
    long long fmin = n_post;
    long long i = pos + lane;
    for (; i + 96 < e; i += 128) {   // 4 independent posting loads and mask gathers in flight
      const int d0 = post_doc[i], d1 = post_doc[i + 32], d2 = post_doc[i + 64], d3 = post_doc[i + 96];
      const int m0 = mask[d0], m1 = mask[d1], m2 = mask[d2], m3 = mask[d3];
      sum += (m0 != 0) + (m1 != 0) + (m2 != 0) + (m3 != 0);
      if (fmin == n_post) fmin = m0 ? i : (m1 ? i + 32 : (m2 ? i + 64 : (m3 ? i + 96 : n_post)));
    }
    for (; i < e; i += 32) {
      const int m = mask[post_doc[i]];
      sum += m != 0;
      if (m && fmin == n_post) fmin = i;
    }
    sum = __reduce_add_sync(0xFFFFFFFFu, sum);
    if (sum > 0) {
      const unsigned int f = __reduce_min_sync(0xFFFFFFFFu, (unsigned int)(fmin - pos));   // offsets < MDF_CHUNK
      if (lane == 0) {
        atomicAdd(&df[t], sum);
        atomicMin(&first[t], (int)(pos + f));
      }
    }
    pos = e;
    if (e == t_end) ++t;
  }
}

}  // namespace cmr

using namespace cmr;

extern "C" int cmr_masked_df(const int64_t* term_ptr, const int32_t* post_doc, int n_terms, int64_t n_postings,
                             const uint8_t* row_mask, int32_t* out_df, int32_t* out_first, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_terms >= 0 && n_postings >= 0 && n_postings < (int64_t)0x7FFFFFFF, "bad index shape");
  CMR_CHECK_ARG(n_terms == 0 || (term_ptr && out_df && out_first), "null pointer argument");
  CMR_CHECK_ARG(n_postings == 0 || (post_doc && row_mask), "null postings / mask");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_terms == 0) return CMR_OK;
  CMR_CUDA(cudaMemsetAsync(out_df, 0, sizeof(int32_t) * (size_t)n_terms, st));
  CMR_CUDA(cudaMemsetAsync(out_first, 0x7F, sizeof(int32_t) * (size_t)n_terms, st));   // 0x7F7F7F7F: above any posting index
  if (n_postings == 0) return CMR_OK;
  const long long warps = (n_postings + MDF_CHUNK - 1) / MDF_CHUNK;
  const long long blocks = (warps * 32 + MDF_THREADS - 1) / MDF_THREADS;
  masked_df_kernel<<<(unsigned int)blocks, MDF_THREADS, 0, st>>>((const long long*)term_ptr, post_doc, n_terms, n_postings,
                                                              row_mask, out_df, out_first);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_gather_rows(const uint16_t* emb, int64_t n_rows, int dim, int64_t row_offset,
                               const int64_t* ids, int n_ids, uint16_t* out_rows, cmr_stream_t stream) {
  CMR_CHECK_ARG(dim > 0 && dim % 8 == 0, "dim must be a multiple of 8");
  CMR_CHECK_ARG(n_ids >= 0 && (n_ids == 0 || (ids && out_rows)), "bad arguments");
  CMR_CHECK_ARG(n_rows == 0 || emb, "null embedding matrix");
  if (n_ids == 0) return CMR_OK;
  gather_rows_kernel<<<n_ids, 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(emb), n_rows, dim / 8, row_offset, (const long long*)ids, n_ids,
      reinterpret_cast<uint4*>(out_rows));
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_mmr_select(const uint16_t* cand_rows, const double* cand_sims, const int64_t* cand_ids,
                              const int32_t* cand_counts, int n_queries, int pool, int dim, int k,
                              double lambda, int64_t* out_ids, double* out_sims, int32_t* out_counts,
                              cmr_stream_t stream) {
  CMR_CHECK_ARG(n_queries > 0, "n_queries must be positive");
  CMR_CHECK_ARG(pool > 0 && pool <= MMR_MAX_POOL, "pool %d out of range (1..%d)", pool, MMR_MAX_POOL);
  CMR_CHECK_ARG(k > 0 && k <= pool, "k %d out of range (1..pool)", k);
  CMR_CHECK_ARG(dim > 0 && dim % 8 == 0, "dim must be a multiple of 8");
  CMR_CHECK_ARG(cand_rows && cand_sims && cand_ids && cand_counts && out_ids && out_sims && out_counts, "null pointer argument");
  // the pool's rows are staged in shared memory when they fit beside the 33 KB of static tables
  const size_t row_bytes = (size_t)pool * dim * 2;
  const int stage = row_bytes <= 160 * 1024;
  if (stage && row_bytes > 12 * 1024)   // per device, and cheap: set on every call rather than remembered
    CMR_CUDA(cudaFuncSetAttribute(mmr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  mmr_select_kernel<<<n_queries, MMR_THREADS, stage ? row_bytes : 0, (cudaStream_t)stream>>>(
      cand_rows, cand_sims, (const long long*)cand_ids, cand_counts, pool, dim, k, lambda,
      (long long*)out_ids, out_sims, out_counts, stage);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_hybrid_fuse(const int64_t* vec_ids, const double* vec_sims, const int32_t* vec_counts, int kv,
                               const int64_t* bm_ids, const double* bm_scores, const int32_t* bm_counts, int kb,
                               int n_queries, double w_vec, double w_bm, int rrf_k, int top_k,
                               int64_t* out_ids, double* out_fused, double* out_vdist, double* out_bm25,
                               int32_t* out_counts, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_queries > 0, "n_queries must be positive");
  CMR_CHECK_ARG(kv >= 0 && kb >= 0 && kv + kb <= FUSE_MAX_ITEMS, "kv + kb must be <= %d", FUSE_MAX_ITEMS);
  CMR_CHECK_ARG(top_k > 0, "top_k must be positive");
  CMR_CHECK_ARG(kv == 0 || (vec_ids && vec_sims && vec_counts), "null vector list");
  CMR_CHECK_ARG(kb == 0 || (bm_ids && bm_scores && bm_counts), "null bm25 list");
  CMR_CHECK_ARG(out_ids && out_fused && out_vdist && out_bm25 && out_counts, "null output");
  hybrid_fuse_kernel<<<n_queries, FUSE_THREADS, 0, (cudaStream_t)stream>>>(
      (const long long*)vec_ids, vec_sims, kv > 0 ? vec_counts : nullptr, kv, (const long long*)bm_ids, bm_scores,
      kb > 0 ? bm_counts : nullptr, kb, w_vec, w_bm, rrf_k, top_k, (long long*)out_ids, out_fused, out_vdist,
      out_bm25, out_counts);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_topk_merge(const double* in_scores, const int64_t* in_ids, const int32_t* in_counts,
                              int n_parts, int n_queries, int k, double* out_scores, int64_t* out_ids,
                              int32_t* out_counts, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_parts > 0 && n_queries > 0 && k > 0, "bad sizes");
  CMR_CHECK_ARG((size_t)n_parts * k * 16 <= 96 * 1024, "n_parts * k too large for one merge (%d x %d)", n_parts, k);
  CMR_CHECK_ARG(in_scores && in_ids && in_counts && out_scores && out_ids && out_counts, "null pointer argument");
  const size_t smem = (size_t)n_parts * k * 16;
  static int attr_dev_mask = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && !(attr_dev_mask & (1 << dev))) {
    CMR_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_dev_mask |= (1 << dev);
  }
  topk_merge_kernel<<<n_queries, MERGE_THREADS, smem, (cudaStream_t)stream>>>(
      in_scores, (const long long*)in_ids, in_counts, n_parts, n_queries, k, out_scores, (long long*)out_ids,
      out_counts);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_rrf_fuse(const int64_t* list_ids, const int32_t* list_counts, int n_lists, int max_len,
                            const double* weights, int rrf_k, int64_t* out_ids, double* out_scores,
                            int32_t* out_count, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_lists >= 1 && max_len >= 1 && (long long)n_lists * max_len <= RRF_MAX_ENTRIES,
                "rrf_fuse: n_lists * max_len must be in 1..%d", RRF_MAX_ENTRIES);
  CMR_CHECK_ARG(list_ids && list_counts && weights && out_ids && out_scores && out_count, "null pointer argument");
  rrf_fuse_kernel<<<1, RRF_THREADS, 0, (cudaStream_t)stream>>>((const long long*)list_ids, list_counts, n_lists,
                                                               max_len, weights, rrf_k, (long long*)out_ids,
                                                               out_scores, out_count);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_filter_mask(const int32_t* field_codes, int64_t n_rows, int n_fields,
                               const int32_t* clause_field, const int32_t* clause_code, int n_clauses,
                               const uint8_t* alive, uint8_t* out_mask, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_rows >= 0 && n_fields >= 0 && n_clauses >= 0 && n_clauses <= 64, "bad filter shape");
  CMR_CHECK_ARG(n_rows == 0 || out_mask, "null output mask");
  CMR_CHECK_ARG(n_clauses == 0 || (field_codes && clause_field && clause_code), "null clause arrays");
  if (n_rows == 0) return CMR_OK;
  long long blocks = (n_rows + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  filter_mask_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(field_codes, n_rows, clause_field, clause_code,
                                                                    n_clauses, alive, out_mask);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" size_t cmr_shard_msg_bytes(int pool, int kb, int dim) {
  if (pool < 0 || kb < 0 || dim < 0 || dim % 8 != 0) return 0;
  return shard_msg_bytes(pool, kb, dim);
}

extern "C" int cmr_shard_pack(const double* dense_scores, const int64_t* dense_ids, const int32_t* dense_counts,
                              const int32_t* dense_flags, int pool, const double* bm_scores, const int64_t* bm_ids,
                              const int32_t* bm_counts, int kb, const uint16_t* emb, int64_t n_rows, int dim,
                              int64_t row_offset, int n_queries, void* msg, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_queries > 0 && pool > 0 && kb >= 0 && dim >= 0 && dim % 8 == 0, "bad shard message shape");
  CMR_CHECK_ARG(dense_scores && dense_ids && dense_counts && msg, "null pointer argument");
  CMR_CHECK_ARG(kb == 0 || (bm_scores && bm_ids && bm_counts), "null BM25 list");
  CMR_CHECK_ARG(dim == 0 || n_rows == 0 || emb, "null embedding matrix");
  CMR_CHECK_ARG(((uintptr_t)msg % 16) == 0, "message buffer must be 16-byte aligned");
  shard_pack_kernel<false><<<n_queries, SHARD_THREADS, 0, (cudaStream_t)stream>>>(
      dense_scores, (const long long*)dense_ids, dense_counts, dense_flags, pool, bm_scores, (const long long*)bm_ids,
      bm_counts, kb, reinterpret_cast<const uint4*>(emb), n_rows, dim, row_offset, (unsigned char*)msg, ShardP2P{});
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

static int launch_shard_merge(const void* gathered, size_t slot_stride, const unsigned int* wait_flags,
                              const unsigned int* wait_state, unsigned long long parity_stride, int* timeout_flag,
                              int n_parts, int n_queries, int pool, int kb, int dim, int msg_dim,
                              const unsigned long long* peer_rows, const long long* peer_row_lo, double* dense_scores,
                              int64_t* dense_ids, int32_t* dense_counts, int32_t* dense_flags, uint16_t* dense_rows,
                              double* bm_scores, int64_t* bm_ids, int32_t* bm_counts, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_parts >= 1 && n_queries > 0 && pool > 0 && kb >= 0 && dim >= 0 && dim % 8 == 0, "bad shard message shape");
  CMR_CHECK_ARG(gathered && dense_scores && dense_ids && dense_counts && dense_flags, "null pointer argument");
  CMR_CHECK_ARG(kb == 0 || (bm_scores && bm_ids && bm_counts), "null BM25 output");
  CMR_CHECK_ARG(n_parts <= SHARD_THREADS, "too many shards");
  const int n = n_parts * (pool > kb ? pool : kb);
  const size_t smem = (size_t)n * 16 + (size_t)pool * 4 + 16;
  CMR_CHECK_ARG(smem <= 48 * 1024, "too many shards x list entries for one merge (%d)", n);
  shard_merge_kernel<<<n_queries, SHARD_THREADS, smem, (cudaStream_t)stream>>>(
      (const unsigned char*)gathered, slot_stride, wait_flags, wait_state, parity_stride, timeout_flag, n_parts,
      n_queries, pool, kb, dim, msg_dim, peer_rows, peer_row_lo, dense_scores, (long long*)dense_ids, dense_counts,
      dense_flags, reinterpret_cast<uint4*>(dense_rows), bm_scores, (long long*)bm_ids, bm_counts);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_shard_merge(const void* gathered, int n_parts, int n_queries, int pool, int kb, int dim,
                               double* dense_scores, int64_t* dense_ids, int32_t* dense_counts, int32_t* dense_flags,
                               uint16_t* dense_rows, double* bm_scores, int64_t* bm_ids, int32_t* bm_counts,
                               cmr_stream_t stream) {
  if (pool < 0 || kb < 0 || dim < 0 || dim % 8 != 0 || n_queries <= 0) {
    set_error("bad shard message shape");
    return CMR_EINVAL;
  }
  return launch_shard_merge(gathered, (size_t)n_queries * shard_msg_bytes(pool, kb, dim), nullptr, nullptr, 0, nullptr,
                            n_parts, n_queries, pool, kb, dim, dim, nullptr, nullptr, dense_scores, dense_ids,
                            dense_counts, dense_flags, dense_rows, bm_scores, bm_ids, bm_counts, stream);
}

extern "C" int cmr_shard_exchange_pack(const double* dense_scores, const int64_t* dense_ids,
                                       const int32_t* dense_counts, const int32_t* dense_flags, int pool,
                                       const double* bm_scores, const int64_t* bm_ids, const int32_t* bm_counts,
                                       int kb, const uint16_t* emb, int64_t n_rows, int dim, int64_t row_offset,
                                       int n_queries, const cmr_shard_p2p* x, cmr_stream_t stream) {
  CMR_CHECK_ARG(n_queries > 0 && pool > 0 && kb >= 0 && dim >= 0 && dim % 8 == 0, "bad shard message shape");
  CMR_CHECK_ARG(dense_scores && dense_ids && dense_counts && x, "null pointer argument");
  CMR_CHECK_ARG(kb == 0 || (bm_scores && bm_ids && bm_counts), "null BM25 list");
  CMR_CHECK_ARG(dim == 0 || n_rows == 0 || emb, "null embedding matrix");
  CMR_CHECK_ARG(x->peer_recv && x->peer_flags && x->state && x->n_parts >= 1 && x->my_rank >= 0 &&
                    x->my_rank < x->n_parts, "incomplete peer-exchange descriptor");
  CMR_CHECK_ARG((x->peer_rows == nullptr) == (x->peer_row_lo == nullptr), "peer_rows and peer_row_lo go together");
  if (x->peer_rows != nullptr) dim = 0;   // pull form: the merge reads rows from their owners, none travel
  CMR_CHECK_ARG((size_t)n_queries * shard_msg_bytes(pool, kb, dim) <= x->slot_stride && x->slot_stride % 16 == 0 &&
                    x->parity_stride >= x->slot_stride * (size_t)x->n_parts && x->parity_stride % 16 == 0,
                "peer receive buffer too small for this message");
  ShardP2P p{(const unsigned long long*)x->peer_recv, (const unsigned long long*)x->peer_flags, (unsigned int*)x->state,
             x->n_parts, x->my_rank, x->slot_stride, x->parity_stride, nullptr, nullptr};
  shard_pack_kernel<true><<<n_queries, SHARD_THREADS, 0, (cudaStream_t)stream>>>(
      dense_scores, (const long long*)dense_ids, dense_counts, dense_flags, pool, bm_scores, (const long long*)bm_ids,
      bm_counts, kb, reinterpret_cast<const uint4*>(emb), n_rows, dim, row_offset, nullptr, p);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_shard_exchange_merge(const void* local_recv, const uint32_t* local_flags, const cmr_shard_p2p* x,
                                        int32_t* timeout_flag, int n_queries, int pool, int kb, int dim,
                                        double* dense_scores, int64_t* dense_ids, int32_t* dense_counts,
                                        int32_t* dense_flags, uint16_t* dense_rows, double* bm_scores,
                                        int64_t* bm_ids, int32_t* bm_counts, cmr_stream_t stream) {
  CMR_CHECK_ARG(local_recv && local_flags && x && x->state && timeout_flag, "null pointer argument");
  CMR_CHECK_ARG((x->peer_rows == nullptr) == (x->peer_row_lo == nullptr), "peer_rows and peer_row_lo go together");
  return launch_shard_merge(local_recv, (size_t)x->slot_stride, local_flags, (const unsigned int*)x->state,
                            x->parity_stride, timeout_flag, x->n_parts, n_queries, pool, kb, dim,
                            x->peer_rows != nullptr ? 0 : dim, (const unsigned long long*)x->peer_rows,
                            (const long long*)x->peer_row_lo, dense_scores, dense_ids, dense_counts, dense_flags,
                            dense_rows, bm_scores, bm_ids, bm_counts, stream);
}
