// bm25.cu -- A2: exact BM25 (Okapi) top-k over a CSR inverted index.
//
// Replaces BM25Okapi.get_scores + the full Python sort inside BM25Store.search
// (reference rag/retrieval/bm25.py:175-212; rank_bm25 semantics restated in
// oracle/np_oracle.py).  Two kernels:
//
//  bm25_tile_kernel      grid (query, tile group), 128 threads, 8 CTAs per SM (small CTAs keep
//                        the per-pass barriers cheap: measured 1.80 ms -> 1.42 ms per 32
//                        queries over 10M documents against 512-thread CTAs).  A CTA walks tiles of
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
//  The batched path (bm25_mma.cu: head-term matrix on the tensor cores, bucketed sparse postings,
//  exact rescoring) hands the queries it could not certify back to these two kernels through
//  `only_flagged`: CTAs of the other queries leave at once.
//
// Algorithmic bytes per query: 4 * df for a sparse token (packed posting), 8 * documents for a
// dense token (its factor column); SURVEY.md section 8(d) counts 8 bytes per posting.
#include <stdlib.h>

#include "bm25_head.cuh"
#include "lex_slice.cuh"
#include "topk.cuh"

namespace cmr {

typedef void (*tile_fn_t)(cmr_lex_index, const int*, const int*, const uint8_t*, KeyD*, const int*);
#ifndef CMR_BM_THREADS
#define CMR_BM_THREADS 128
#endif
#ifndef CMR_BM_MINCTAS
#define CMR_BM_MINCTAS 8
#endif
#ifndef CMR_BM_SWEEP_U
#define CMR_BM_SWEEP_U 8
#endif
#ifndef CMR_BM_RUN
#define CMR_BM_RUN 2
#endif
constexpr int BM_THREADS = CMR_BM_THREADS;
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

constexpr int BM_MAX_LISTS = 2048;  // tile groups (= sorted lists handed to finalize) per query
constexpr int BM_MAXQ = 64;  // query tokens staged per chunk
constexpr int BM_U = 8;      // postings in flight per thread
constexpr int BM_RUN = CMR_BM_RUN;  static_assert(BM_RUN == 1 || BM_RUN == 2, "sweep is specialised for runs of 1 or 2");
//    // consecutive dense tokens fused into one sweep
constexpr int BM_SWEEP_U = CMR_BM_SWEEP_U;  // documents per thread in flight in a dense sweep

__device__ __forceinline__ u32 ldg_stream_u32(const u32* p) {
  u32 r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// Dense sweep over a FULL tile whose size is a multiple of BM_SWEEP_U * BM_THREADS: no
// bounds or run predicates in the loop.  RUN consecutive dense tokens (columns c0, c1 with
// weights w0, w1) share one read-modify-write of the accumulators; FRESH: the accumulators
// have not been written for this tile yet, start from 0.0 instead of reading them.
template <int RUN, bool FRESH>
__device__ __forceinline__ void bm25_sweep_full(double* __restrict__ acc, int tile_docs,
                                                const double* __restrict__ c0,
                                                const double* __restrict__ c1, double w0, double w1,
                                                int tid) {
  for (int i0 = tid; i0 < tile_docs; i0 += BM_SWEEP_U * BM_THREADS) {
    double f0[BM_SWEEP_U], f1[BM_SWEEP_U];
#pragma unroll
    for (int u = 0; u < BM_SWEEP_U; ++u) {
      f0[u] = ldg_stream_f64(c0 + i0 + u * BM_THREADS);
      if (RUN == 2) f1[u] = ldg_stream_f64(c1 + i0 + u * BM_THREADS);
    }
#pragma unroll
    for (int u = 0; u < BM_SWEEP_U; ++u) {
      double a = FRESH ? 0.0 : acc[i0 + u * BM_THREADS];
      a = __dadd_rn(a, __dmul_rn(w0, f0[u]));  // no fma: rank_bm25 rounds the product first
      if (RUN == 2) a = __dadd_rn(a, __dmul_rn(w1, f1[u]));
      acc[i0 + u * BM_THREADS] = a;
    }
  }
}

// PACKED: 4-byte postings (code << 16 | tile-local doc) + float64 factor table;
// otherwise int32 doc + float64 factor per posting.
template <int KPL, bool PACKED>
__global__ void __launch_bounds__(BM_THREADS, CMR_BM_MINCTAS)
bm25_tile_kernel(cmr_lex_index ix, const int* __restrict__ q_terms, const int* __restrict__ q_ptr,
                 const uint8_t* __restrict__ row_mask, KeyD* __restrict__ part,
                 const int* __restrict__ only_flagged) {
  constexpr int KP = 32 * KPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);             // [tile_docs]
  KeyD* s_lists = reinterpret_cast<KeyD*>(acc + ix.tile_docs);   // [warps][KP]
  KeyD* s_out = s_lists + BM_WARPS * KP;                         // [KP]
  double* s_thr = reinterpret_cast<double*>(s_out + KP);         // [warps]
  double* s_w = s_thr + BM_WARPS;                                // [MAXQ] idf of each staged token
  long long* s_base = reinterpret_cast<long long*>(s_w + BM_MAXQ);   // [MAXQ] term_ptr
  const uint32_t** s_skip = reinterpret_cast<const uint32_t**>(s_base + BM_MAXQ);  // [MAXQ] skip row
  long long* s_lo = reinterpret_cast<long long*>(s_skip + BM_MAXQ);  // [2][MAXQ] slice bounds
  long long* s_hi = s_lo + 2 * BM_MAXQ;
  double* s_seed_buf = reinterpret_cast<double*>(s_hi + 2 * BM_MAXQ);  // [128] group maxima (seeding)
  int* s_slot = reinterpret_cast<int*>(s_seed_buf + 128);              // [MAXQ] dense column of the token or -1

  const int b = blockIdx.x;
  if (only_flagged != nullptr && only_flagged[b] == 0) return;  // served by the batched path (CTA-uniform exit)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < BM_WARPS * KP; i += BM_THREADS) key_clear(s_lists[i]);
  if (tid < BM_WARPS) s_thr[tid] = -INFINITY;
  const int qlo = q_ptr[b], qhi = q_ptr[b + 1];
  KeyD* w_list = s_lists + warp * KP;
  volatile double* w_thr = s_thr + warp;
  const int first_tile = blockIdx.y;

  // Query tokens are processed in chunks of BM_MAXQ (one chunk for any normal query).
  // With more than one chunk the accumulators must persist across chunks, so the tile
  // loop is the outer loop and chunks the inner one.
  //
  // Single-chunk queries (the normal case) keep a token's constants -- posting base, idf,
  // dense column -- in the registers of thread `tid` (< 64) for the whole kernel; per tile
  // only the two skip-table entries of SPARSE tokens are needed, and those of the next tile
  // are loaded while this tile is processed, so staging adds no dependent load to the
  // per-tile chain.
  const bool one_chunk = (qhi - qlo) <= BM_MAXQ;
  long long tk_base = 0;
  int tk_term = -1, tk_row = -1;   // sparse token of this thread: its term and its skip-table row (-1: bisection)
  u32 pre0 = 0, pre1 = 0;  // slice (relative to tk_base) of the tile about to be processed
  if (one_chunk && tid < qhi - qlo) {
    const int t = q_terms[qlo + tid];
    double w = 0.0;
    int slot = -1;
    if (t >= 0 && t < ix.n_terms) {
      tk_base = ix.term_ptr[t];
      w = ix.idf[t];
      if (ix.dense_slot != nullptr) slot = ix.dense_slot[t];
      if (slot < 0) {
        tk_term = t;
        tk_row = lex_skip_row(ix, t);
        if (first_tile < ix.n_tiles) lex_slice(ix, t, tk_row, first_tile, &pre0, &pre1);
      }
    }
    s_w[tid] = w;
    s_slot[tid] = slot;
  }

  for (int tile = first_tile; tile < ix.n_tiles; tile += gridDim.y) {
    const long long tile_lo = (long long)tile * ix.tile_docs;
    const long long rem = ix.n_docs - tile_lo;
    const int n_here = rem < ix.tile_docs ? (int)rem : ix.tile_docs;
    bool fresh = true;  // accumulators of this tile not written yet (a leading dense sweep starts from 0.0)
    const bool full_tile = n_here == ix.tile_docs && (ix.tile_docs % (BM_SWEEP_U * BM_THREADS)) == 0;
    for (int i = n_here + tid; i < ix.tile_docs; i += BM_THREADS) acc[i] = 0.0;  // ragged last tile

    for (int c0 = qlo; c0 < qhi; c0 += BM_MAXQ) {
      const int m = (qhi - c0) < BM_MAXQ ? (qhi - c0) : BM_MAXQ;
      // A later chunk of a long query re-uses the staging arrays the previous chunk's passes read.
      // (The first chunk needs no barrier: the tile loop ends with one, and before the first tile
      // nobody has read the staging arrays yet.)
      if (c0 != qlo) __syncthreads();
      if (one_chunk) {
        if (tid < m) {
          s_lo[tid] = tk_base + pre0;   // dense / unknown tokens: an empty slice (0, 0)
          s_hi[tid] = tk_base + pre1;
          const int next = tile + (int)gridDim.y;
          if (tk_term >= 0 && next < ix.n_tiles)   // in flight during this tile's passes
            lex_slice(ix, tk_term, tk_row, next, &pre0, &pre1);
        }
      } else if (tid < m) {
        const int t = q_terms[c0 + tid];
        long long lo = 0, hi = 0;
        double w = 0.0;
        int slot = -1;
        if (t >= 0 && t < ix.n_terms) {  // unknown token: empty slice
          const long long base = ix.term_ptr[t];
          u32 a = 0, z = 0;
          lex_slice(ix, t, lex_skip_row(ix, t), tile, &a, &z);
          lo = base + a;
          hi = base + z;
          w = ix.idf[t];
          if (ix.dense_slot != nullptr) slot = ix.dense_slot[t];
        }
        s_w[tid] = w;
        s_lo[tid] = lo;
        s_hi[tid] = hi;
        s_slot[tid] = slot;
      }
      __syncthreads();

      // token passes, strictly in query order (float64 addition order = rank_bm25's)
      int j = 0;
      while (j < m) {
        if (s_slot[j] >= 0) {
          // ---- dense term(s): owner-computes sweep over the tile ------------------------
          // Up to BM_RUN consecutive dense tokens share one read-modify-write of the
          // accumulators.  Thread t owns documents t, t+512, ...: coalesced 8-byte loads of
          // the terms' factor columns (0.0 where the term is absent: x + 0.0 == x), conflict
          // free shared-memory accesses, no atomics.
          int run = 1;
          while (run < BM_RUN && j + run < m && s_slot[j + run] >= 0) ++run;
          if (full_tile) {
            const double* c0p = ix.dense_imp + (size_t)s_slot[j] * ix.n_docs + tile_lo;
            const double w0 = s_w[j];
            if (run == 2) {
              const double* c1p = ix.dense_imp + (size_t)s_slot[j + 1] * ix.n_docs + tile_lo;
              const double w1 = s_w[j + 1];
              if (fresh) bm25_sweep_full<2, true>(acc, ix.tile_docs, c0p, c1p, w0, w1, tid);
              else bm25_sweep_full<2, false>(acc, ix.tile_docs, c0p, c1p, w0, w1, tid);
            } else {
              if (fresh) bm25_sweep_full<1, true>(acc, ix.tile_docs, c0p, c0p, w0, w0, tid);
              else bm25_sweep_full<1, false>(acc, ix.tile_docs, c0p, c0p, w0, w0, tid);
            }
          } else {
            // ragged last tile / odd tile size: predicated version
            const double* col[BM_RUN];
            double w[BM_RUN];
#pragma unroll
            for (int r = 0; r < BM_RUN; ++r) {
              const int jj = j + (r < run ? r : 0);
              col[r] = ix.dense_imp + (size_t)s_slot[jj] * ix.n_docs + tile_lo;
              w[r] = s_w[jj];
            }
            for (int i = tid; i < n_here; i += BM_THREADS) {
              double a = fresh ? 0.0 : acc[i];
#pragma unroll
              for (int r = 0; r < BM_RUN; ++r)
                if (r < run) a = __dadd_rn(a, __dmul_rn(w[r], ldg_stream_f64(col[r] + i)));
              acc[i] = a;
            }
          }
          fresh = false;
          j += run;
          __syncthreads();
          continue;
        }
        // ---- sparse term: scatter its slice of postings into the accumulators -----------
        const long long lo = s_lo[j];
        const int n = (int)(s_hi[j] - lo);  // <= tile_docs
        if (n > 0) {
          if (fresh) {
            for (int i = tid; i < ix.tile_docs; i += BM_THREADS) acc[i] = 0.0;
            fresh = false;
            __syncthreads();
          }
          const double w = s_w[j];
          const u32* pp = ix.post_pack + lo;
          const int* pd = ix.post_doc + lo;
          const double* pi = ix.post_imp + lo;
          int i0 = 0;
          for (; i0 + BM_U * BM_THREADS <= n; i0 += BM_U * BM_THREADS) {  // full chunks: no predicates
            u32 k2[BM_U];
            double v2[BM_U];
#pragma unroll
            for (int u = 0; u < BM_U; ++u) {
              const int i = i0 + tid + u * BM_THREADS;
              if (PACKED) {
                k2[u] = ldg_stream_u32(pp + i);
              } else {
                k2[u] = (u32)(ldg_stream_i32(pd + i) - (int)tile_lo);
                v2[u] = ldg_stream_f64(pi + i);
              }
            }
#pragma unroll
            for (int u = 0; u < BM_U; ++u) {
              const u32 loc = PACKED ? (k2[u] & 0xFFFFu) : k2[u];
              const double imp = PACKED ? __ldg(ix.imp_table + (k2[u] >> 16)) : v2[u];
              acc[loc] = __dadd_rn(acc[loc], __dmul_rn(w, imp));  // documents are unique within a list
            }
          }
          // tail (for a sparse term usually the whole slice): warps past the end skip at once
          for (int i = i0 + tid; i < n; i += BM_THREADS) {
            u32 loc;
            double imp;
            if (PACKED) {
              const u32 pk = ldg_stream_u32(pp + i);
              loc = pk & 0xFFFFu;
              imp = __ldg(ix.imp_table + (pk >> 16));
            } else {
              loc = (u32)(ldg_stream_i32(pd + i) - (int)tile_lo);
              imp = ldg_stream_f64(pi + i);
            }
            acc[loc] = __dadd_rn(acc[loc], __dmul_rn(w, imp));
          }
          __syncthreads();  // a document may appear in the next token's list too
        }
        ++j;
      }
    }
    if (fresh) {  // no token touched this tile: every document scores 0.0 (otherwise the last pass ended with a barrier)
      for (int i = tid; i < ix.tile_docs; i += BM_THREADS) acc[i] = 0.0;
      __syncthreads();
    }

    // Threshold seeding (first tile of this CTA only).  Split the tile into BM_THREADS/4
    // groups of documents and take each group's best score: the KP-th largest of
    // those 128 maxima is reached by at least KP documents, so nothing below it can
    // be in this CTA's top KP.  Every warp starts from the largest double below
    // that value (ties are admitted until its own list is full) instead of -inf,
    // which removes almost all of the list-filling inserts.
    if (tile == first_tile) {
      double m = -INFINITY;
      for (int i = tid; i < n_here; i += BM_THREADS) {
        bool ok = true;
        if (row_mask != nullptr) ok = row_mask[tile_lo + i] != 0;
        if (ok) m = fmax(m, acc[i]);
      }
      m = fmax(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
      m = fmax(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
      if ((tid & 3) == 0) s_seed_buf[tid >> 2] = m;
      __syncthreads();
      constexpr int SEED_GROUPS = BM_THREADS / 4;  // <= 128 slots in s_seed_buf
      static_assert(SEED_GROUPS <= 128, "s_seed_buf holds 128 group maxima");
      if (tid < SEED_GROUPS) {
        const double mine = s_seed_buf[tid];
        int cnt = 0;
        for (int j = 0; j < SEED_GROUPS; ++j) {
          const double o = s_seed_buf[j];
          cnt += (o > mine) || (o == mine && j < tid);
        }
        if (cnt == KP - 1 && mine > -INFINITY) {
          // largest double strictly below `mine` (mine is finite)
          const long long bits = __double_as_longlong(mine);
          double below;
          if (mine > 0.0) below = __longlong_as_double(bits - 1);
          else if (mine < 0.0) below = __longlong_as_double(bits + 1);
          else below = -4.9406564584124654e-324;  // -denorm_min (mine is +-0.0)
          for (int w = 0; w < BM_WARPS; ++w) s_thr[w] = below;
        }
      }
      __syncthreads();
    }

    // per-warp selection over the tile's accumulators.  A warp owns a slice and
    // visits it 128 documents at a time: lane l looks at documents l, l+32, l+64,
    // l+96 of the chunk, so hits are inserted in ascending document order and the
    // strict '>' admission test is exact (a tie with the list's last entry always
    // loses the id tie-break).  Fast path: 4 LDS.64 + 4 compares per 128 documents.
    {
      const int per_warp = ix.tile_docs / BM_WARPS;  // multiple of 32
      const int w_lo = warp * per_warp;
      const int w_hi = w_lo + per_warp;
      for (int c = w_lo; c < w_hi; c += 128) {
        double v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = c + e * 32 + lane;
          v[e] = (i < w_hi) ? acc[i] : -INFINITY;
        }
        const double thr = *w_thr;
        const bool any_hit = (v[0] > thr) | (v[1] > thr) | (v[2] > thr) | (v[3] > thr);
        if (!__any_sync(0xFFFFFFFFu, any_hit)) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = c + e * 32 + lane;
          bool ok = (i < w_hi) && (i < n_here) && (v[e] > *w_thr);
          if (ok && row_mask != nullptr) ok = row_mask[tile_lo + i] != 0;
          unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
          while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            KeyD key;
            key.s = __shfl_sync(0xFFFFFFFFu, v[e], src);
            key.id = (u32)(tile_lo + c + e * 32 + src);
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
                     int* __restrict__ out_counts, int* __restrict__ out_flags,
                     const int* __restrict__ only_flagged) {
  constexpr int KP = 32 * KPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // only_flagged aliases out_flags: every thread reads it here, thread 0 clears it after the barriers below
  if (only_flagged != nullptr && only_flagged[blockIdx.x] == 0) return;
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
  bool packed;
  tile_fn_t fn;
  int kpl;
  int grid_y;  // tile groups (lists per query)
  size_t smem_tile, smem_fin;
};

static int check_index(const cmr_lex_index* ix) {
  CMR_CHECK_ARG(ix != nullptr, "null index");
  CMR_CHECK_ARG(ix->n_docs >= 0 && ix->n_docs < 0xFFFFFFFFll, "n_docs out of range");
  CMR_CHECK_ARG(ix->tile_docs >= 512 && ix->tile_docs % 512 == 0 && ix->tile_docs <= 65536,
                "tile_docs must be a multiple of 512 in [512, 65536]");
  CMR_CHECK_ARG(ix->n_terms == 0 || (ix->post_pack && ix->imp_table) || (ix->post_doc && ix->post_imp) || ix->term_ptr,
                "index needs packed or wide postings");
  CMR_CHECK_ARG(ix->n_tiles >= 1 && (long long)ix->n_tiles * ix->tile_docs >= ix->n_docs, "n_tiles inconsistent with n_docs/tile_docs");
  CMR_CHECK_ARG(ix->n_terms >= 0, "n_terms negative");
  CMR_CHECK_ARG(ix->n_terms == 0 || (ix->term_ptr && ix->idf && (ix->tile_skip || ix->skip_row)), "null index arrays");
  CMR_CHECK_ARG(ix->skip_row == nullptr || ix->post_doc != nullptr, "skip_row needs post_doc (bisection of the short lists)");
  CMR_CHECK_ARG(ix->n_dense >= 0 && (ix->n_dense == 0 || (ix->dense_imp && ix->dense_slot)), "dense columns inconsistent");
  return CMR_OK;
}

// rerun: the plan serves the few queries the head path flagged (every CTA of the other queries
// leaves at once), so the wave is sized for 4 queries, with the launch kept below ~4k CTAs (an
// empty launch of 9 500 CTAs costs 7 us, one of 4 000 costs 3)
static int make_plan(const cmr_lex_index& ix, int n_queries, bool rerun, int k, Bm25Plan* p) {
  p->kpl = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
  const int kp = 32 * p->kpl;
  p->smem_tile = (size_t)ix.tile_docs * 8 + (size_t)BM_WARPS * kp * 16 + (size_t)kp * 16 + BM_WARPS * 8 +
                 (size_t)BM_MAXQ * (8 + 8 + 8 + 4 * 8) + 128 * 8 + (size_t)BM_MAXQ * 4 + 16;
  p->packed = ix.post_pack != nullptr && ix.imp_table != nullptr;
  if (p->smem_tile > 220 * 1024) {
    set_error("bm25 tile shared memory %zu too large: lower tile_docs", p->smem_tile);
    return CMR_EUNSUPPORTED;
  }
  const int sms = sm_count();
  if (sms <= 0) return CMR_ECUDA;
  tile_fn_t fn;
  if (p->packed) fn = p->kpl == 1 ? bm25_tile_kernel<1, true> : (p->kpl == 2 ? bm25_tile_kernel<2, true> : bm25_tile_kernel<4, true>);
  else fn = p->kpl == 1 ? bm25_tile_kernel<1, false> : (p->kpl == 2 ? bm25_tile_kernel<2, false> : bm25_tile_kernel<4, false>);
  p->fn = fn;
  struct Occ { tile_fn_t fn; size_t smem; int dev; int per_sm; };
  static Occ cache[32];
  static int n_cache = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 0;
  for (int i = 0; i < n_cache; ++i)
    if (cache[i].fn == fn && cache[i].smem == p->smem_tile && cache[i].dev == dev) per_sm = cache[i].per_sm;
  if (per_sm == 0) {
    // largest size any plan can ask for (per-kernel attribute: never lower it for a later shape)
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(bm25_tile)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, BM_THREADS, p->smem_tile);
    if (e != cudaSuccess || per_sm <= 0) {
      set_error("occupancy query failed for bm25_tile (smem=%zu)", p->smem_tile);
      return CMR_ECUDA;
    }
    if (n_cache < 32) cache[n_cache++] = Occ{fn, p->smem_tile, dev, per_sm};
  }
  // CMR_BM25_CTAS_PER_SM: plan the single wave for fewer CTAs per SM than the kernel could hold
  // alone -- set when the kernel shares the SMs with the dense scan (engine overlap)
  {
    const char* e = getenv("CMR_BM25_CTAS_PER_SM");  // read per call: cheap, and a process may change it
    const int cap_per_sm = e ? atoi(e) : 0;
    if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
  }
  const long long resident = (long long)sms * per_sm;
  long long gy = resident / n_queries;  // one wave: every CTA resident
  if (rerun) {
    const long long few = resident / (n_queries < 4 ? n_queries : 4), room = 4096 / n_queries;
    const long long wide = few < room ? few : room;
    if (wide > gy) gy = wide;
  }
  if (gy < 1) gy = 1;
  if (gy > ix.n_tiles) gy = ix.n_tiles;
  // the finalize kernel ranks the list heads of a query against each other: keep the number
  // of lists per query bounded (matters for one or two queries, where gy would be ~1000)
  static long long max_gy = -1;
  if (max_gy < 0) {
    const char* e = getenv("CMR_BM25_MAX_LISTS");
    max_gy = e ? atoll(e) : BM_MAX_LISTS;
    if (max_gy < 1) max_gy = BM_MAX_LISTS;
  }
  if (gy > max_gy) gy = max_gy;
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
                       KeyD* part, const int* only_flagged, cudaStream_t st) {
  static int attr_dev_mask = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_dev_mask & (1 << dev))) {
    cudaError_t e = cudaFuncSetAttribute(bm25_finalize_kernel<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(bm25_finalize)");
    attr_dev_mask |= (1 << dev);
  }
  dim3 grid(n_queries, p.grid_y);
  p.fn<<<grid, BM_THREADS, p.smem_tile, st>>>(ix, q_terms, q_ptr, row_mask, part, only_flagged);
  bm25_finalize_kernel<KPL><<<n_queries, BMF_THREADS, p.smem_fin, st>>>(part, p.grid_y, row_offset, k, out_scores,
                                                                       out_ids, out_counts, out_flags, only_flagged);
  return CMR_OK;
}

// Shared with the exact dense scan (dense.cu): select the k best (float64 score, id) keys over
// n_lists sorted lists per query.  kpl in {1, 2, 4}.
int launch_keyd_finalize(const KeyD* part, int n_lists, int n_queries, int kpl, long long row_offset, int k,
                         double* out_scores, long long* out_ids, int* out_counts, int* out_flags,
                         cudaStream_t st) {
  const int kp = 32 * kpl;
  const int cap = kp * kp < 4096 ? kp * kp : 4096;
  const size_t smem = (size_t)(n_lists + 1) * 16 + (size_t)cap * 16 + (size_t)kp * 16 + 16;
  if (smem > 220 * 1024) {
    set_error("finalize shared memory %zu too large", smem);
    return CMR_EUNSUPPORTED;
  }
  static int attr_dev_mask[3] = {0, 0, 0};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = kpl == 1 ? 0 : (kpl == 2 ? 1 : 2);
  if (!(attr_dev_mask[slot] & (1 << dev))) {
    cudaError_t e = kpl == 1 ? cudaFuncSetAttribute(bm25_finalize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)
                  : kpl == 2 ? cudaFuncSetAttribute(bm25_finalize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)
                             : cudaFuncSetAttribute(bm25_finalize_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(keyd_finalize)");
    attr_dev_mask[slot] |= (1 << dev);
  }
  if (kpl == 1) bm25_finalize_kernel<1><<<n_queries, BMF_THREADS, smem, st>>>(part, n_lists, row_offset, k, out_scores, out_ids, out_counts, out_flags, nullptr);
  else if (kpl == 2) bm25_finalize_kernel<2><<<n_queries, BMF_THREADS, smem, st>>>(part, n_lists, row_offset, k, out_scores, out_ids, out_counts, out_flags, nullptr);
  else bm25_finalize_kernel<4><<<n_queries, BMF_THREADS, smem, st>>>(part, n_lists, row_offset, k, out_scores, out_ids, out_counts, out_flags, nullptr);
  return CMR_OK;
}

// the exact kernels over every query (only_flagged == nullptr) or over the flagged ones
static int run_exact(const cmr_lex_index& ix, const Bm25Plan& p, const int* q_terms, const int* q_ptr, int n_queries,
                     int k, const uint8_t* row_mask, long long row_offset, double* out_scores, long long* out_ids,
                     int* out_counts, int* out_flags, KeyD* part, const int* only_flagged, cudaStream_t st) {
  switch (p.kpl) {
    case 1: return launch_bm25<1>(ix, p, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, out_ids, out_counts, out_flags, part, only_flagged, st);
    case 2: return launch_bm25<2>(ix, p, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, out_ids, out_counts, out_flags, part, only_flagged, st);
    default: return launch_bm25<4>(ix, p, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, out_ids, out_counts, out_flags, part, only_flagged, st);
  }
}

static int head_min_queries() {
  const char* e = getenv("CMR_BM25_HEAD_MIN_QUERIES");  // read per call: cheap, and a process may change it
  const int n = e ? atoi(e) : 8;
  return n < 1 ? 1 : n;
}

// which kernels serve this call: 0 = exact only, 1 = head path + exact re-run of flagged queries,
// 2 = head path alone
static int resolve_algo(const cmr_lex_index& ix, int n_queries, int k, bool has_mask, int algo, int* mode) {
  const bool can = bm25_head_eligible(ix, k, has_mask);
  switch (algo) {
    case CMR_BM25_AUTO: *mode = (can && n_queries >= head_min_queries()) ? 1 : 0; return CMR_OK;
    case CMR_BM25_EXACT: *mode = 0; return CMR_OK;
    case CMR_BM25_HEAD:
    case CMR_BM25_HEAD_NOFALLBACK:
      if (!can) {
        set_error("CMR_BM25_HEAD needs head_mat + packed postings, no row_mask, >= 128 documents, tile_docs <= 2048");
        return CMR_EUNSUPPORTED;
      }
      *mode = algo == CMR_BM25_HEAD ? 1 : 2;
      return CMR_OK;
    default: set_error("unknown bm25 algo %d", algo); return CMR_EINVAL;
  }
}

// the exact kernels' share of the workspace comes first, the head path's after it
static size_t exact_ws_bytes(const Bm25Plan& p, int n_queries) {
  return ((size_t)n_queries * p.grid_y * (32 * p.kpl) * sizeof(KeyD) + 255) / 256 * 256;
}

}  // namespace cmr

using namespace cmr;

extern "C" size_t cmr_bm25_workspace_bytes(const cmr_lex_index* ix, int n_queries, int k) {
  if (check_index(ix) != CMR_OK || n_queries <= 0 || k <= 0 || k > CMR_MAX_K) return 0;
  Bm25Plan p, pf;
  if (make_plan(*ix, n_queries, false, k, &p) != CMR_OK) return 0;
  size_t need = exact_ws_bytes(p, n_queries);
  if (bm25_head_eligible(*ix, k, false)) {
    if (make_plan(*ix, n_queries, true, k, &pf) != CMR_OK) return 0;
    const size_t fb = exact_ws_bytes(pf, n_queries);
    if (fb > need) need = fb;
    need += bm25_head_workspace_bytes(*ix, n_queries, k);
  }
  return need;
}

extern "C" int cmr_bm25_topk_ex(const cmr_lex_index* ix, const int32_t* q_terms, const int32_t* q_ptr,
                                int n_queries, int k, const uint8_t* row_mask, int64_t row_offset,
                                double* out_scores, int64_t* out_ids, int32_t* out_counts,
                                int32_t* out_flags, void* workspace, size_t workspace_bytes,
                                cmr_stream_t stream, int algo) {
  int rc = check_index(ix);
  if (rc != CMR_OK) return rc;
  CMR_CHECK_ARG(n_queries > 0 && n_queries <= 1 << 20, "n_queries %d out of range", n_queries);
  CMR_CHECK_ARG(k > 0 && k <= CMR_MAX_K, "k %d out of range (1..%d)", k, CMR_MAX_K);
  CMR_CHECK_ARG(q_ptr && out_scores && out_ids && out_counts && out_flags, "null pointer argument");
  int mode = 0;
  rc = resolve_algo(*ix, n_queries, k, row_mask != nullptr, algo, &mode);
  if (rc != CMR_OK) return rc;
  // the exact kernels plan one wave of CTAs over the queries they serve: all of them, or (after
  // the head path) the few it flagged
  Bm25Plan p;
  rc = make_plan(*ix, n_queries, mode != 0, k, &p);
  if (rc != CMR_OK) return rc;
  const size_t exact_bytes = exact_ws_bytes(p, n_queries);
  const size_t need = exact_bytes + (mode != 0 ? bm25_head_workspace_bytes(*ix, n_queries, k) : 0);
  if (!workspace || workspace_bytes < need) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return CMR_EWORKSPACE;
  }
  CMR_CHECK_ARG(((uintptr_t)workspace % 16) == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  KeyD* part = (KeyD*)workspace;
  long long* ids = (long long*)out_ids;
  const int* only_flagged = nullptr;
  if (mode != 0) {
    rc = bm25_head_topk(*ix, q_terms, q_ptr, n_queries, k, row_offset, out_scores, ids, out_counts, out_flags,
                        (unsigned char*)workspace + exact_bytes, workspace_bytes - exact_bytes, st);
    if (rc != CMR_OK) return rc;
    if (mode == 2) {
      CMR_CUDA(cudaGetLastError());
      return CMR_OK;
    }
    only_flagged = out_flags;
  }
  rc = run_exact(*ix, p, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, ids, out_counts, out_flags,
                 part, only_flagged, st);
  if (rc != CMR_OK) return rc;
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_bm25_topk(const cmr_lex_index* ix, const int32_t* q_terms, const int32_t* q_ptr,
                             int n_queries, int k, const uint8_t* row_mask, int64_t row_offset,
                             double* out_scores, int64_t* out_ids, int32_t* out_counts,
                             int32_t* out_flags, void* workspace, size_t workspace_bytes,
                             cmr_stream_t stream) {
  return cmr_bm25_topk_ex(ix, q_terms, q_ptr, n_queries, k, row_mask, row_offset, out_scores, out_ids, out_counts,
                          out_flags, workspace, workspace_bytes, stream, CMR_BM25_AUTO);
}
