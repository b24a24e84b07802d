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
//  bm25_batch_kernel     (opt-in, CMR_BM25_BATCH=1) the tile-parallel form for batches: one
//                        persistent CTA per SM owns document tiles, all queries of a chunk visit
//                        a slice while its dense columns sit in shared memory (bulk copies +
//                        mbarriers), accumulators in registers.  Bit-identical output; at parity
//                        with the tile kernel on B200, hence not the default.
//
// Algorithmic bytes per query: 4 * df for a sparse token (packed posting), 8 * documents for a
// dense token (its factor column); SURVEY.md section 8(d) counts 8 bytes per posting.
#include <stdlib.h>

#include "topk.cuh"

namespace cmr {

typedef void (*tile_fn_t)(cmr_lex_index, const int*, const int*, const uint8_t*, KeyD*, int);
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
                 const uint8_t* __restrict__ row_mask, KeyD* __restrict__ part, int min_tokens) {
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
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < BM_WARPS * KP; i += BM_THREADS) key_clear(s_lists[i]);
  if (tid < BM_WARPS) s_thr[tid] = -INFINITY;
  const int qlo = q_ptr[b], qhi = q_ptr[b + 1];
  if (qhi - qlo <= min_tokens) return;  // the batch kernel has served this query (CTA-uniform exit)
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
  const uint32_t* tk_skip = nullptr;
  u32 pre0 = 0, pre1 = 0;  // skip entries of the tile about to be processed
  if (one_chunk && tid < qhi - qlo) {
    const int t = q_terms[qlo + tid];
    double w = 0.0;
    int slot = -1;
    if (t >= 0 && t < ix.n_terms) {
      tk_base = ix.term_ptr[t];
      w = ix.idf[t];
      if (ix.dense_slot != nullptr) slot = ix.dense_slot[t];
      if (slot < 0) {
        tk_skip = ix.tile_skip + (size_t)t * (ix.n_tiles + 1);
        if (first_tile < ix.n_tiles) {
          pre0 = tk_skip[first_tile];
          pre1 = tk_skip[first_tile + 1];
        }
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
          if (tk_skip != nullptr && next < ix.n_tiles) {  // in flight during this tile's passes
            pre0 = tk_skip[next];
            pre1 = tk_skip[next + 1];
          }
        }
      } else if (tid < m) {
        const int t = q_terms[c0 + tid];
        long long lo = 0, hi = 0;
        double w = 0.0;
        int slot = -1;
        if (t >= 0 && t < ix.n_terms) {  // unknown token: empty slice
          const long long base = ix.term_ptr[t];
          const uint32_t* sk = ix.tile_skip + (size_t)t * (ix.n_tiles + 1) + tile;
          lo = base + sk[0];
          hi = base + sk[1];
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


// ---------------------------------------------------------------------------------------
// bm25_batch_kernel -- the batched form (>= BB_MIN_QUERIES queries, k <= 32, packed postings).
//
// The tile kernel above re-reads a dense term's factor column from L2 once per query that
// contains the term: 32 queries x ~3 dense tokens x 80 MB at 10M documents = 8.8 GB of
// L2->SM traffic for 2.9 GB of distinct column data, and every token pass ends in a CTA
// barrier.  Here a CTA owns document tiles and ALL queries of a chunk (<= 32) visit a
// tile while it is on chip:
//   * producer warp: one bulk copy (cp.async.bulk + mbarrier complete_tx) per distinct dense
//     column used by the chunk stages a 256-document slice (2 KB) of it in shared memory,
//     double buffered -- every column byte leaves HBM/L2 once per chunk, not once per query;
//   * 16 consumer warps, each owning 2 queries of the chunk (paired heavy + light by an
//     estimate of their per-slice cost): lane l keeps the float64 accumulators of documents
//     l, l+32, ... (8 per lane) in REGISTERS and walks the query's tokens in order.  A dense
//     token is 8 conflict-free LDS.64 + DMUL + DADD; no barrier, no shared-memory
//     read-modify-write.  A sparse token with postings in the slice spills the 8 accumulators
//     to a warp-private 2 KB strip, scatters the postings into it (__syncwarp only), and the
//     next dense token reloads.  Token constants live in the registers of lane j (token j)
//     and are broadcast with shuffles;
//   * the postings a query needs for a 2048-document tile (~150) are copied into a
//     warp-private arena with cp.async one tile ahead (skip-table entries two tiles ahead),
//     so the token loop never waits for a posting to arrive from HBM;
//   * the scores are compared with the query's admission threshold straight from registers;
//     hits go to the (CTA, query) sorted list in shared memory.  Additions happen in query
//     token order for every document, so the float64 scores equal the tile kernel's bit for bit.
// Queries longer than BB_MAXT tokens are left to the tile kernel (launched afterwards in
// its "long queries only" mode).
// ---------------------------------------------------------------------------------------
constexpr int BB_SUB = 256;                 // documents per staged slice
constexpr int BB_R = BB_SUB / 32;           // accumulators per lane
constexpr int BB_STAGES = 2;
#ifndef CMR_BB_NC
#define CMR_BB_NC 24
#endif
constexpr int BB_NC = CMR_BB_NC;            // dense columns staged per slice (others are read from L2)
constexpr int BB_CWARPS = 16;               // consumer warps
constexpr int BB_QW = 2;                    // queries per consumer warp
constexpr int BB_QB = BB_CWARPS * BB_QW;    // queries per chunk
constexpr int BB_THREADS = (BB_CWARPS + 1) * 32;
constexpr int BB_MAXT = 16;                 // tokens per query served here (one per lane, 16-byte program entry each)
constexpr int BB_KP = 32;
constexpr int BB_ARENA = 192;               // postings staged per (query, tile); the rest is read from global
constexpr int BB_MAX_DENSE = 512;           // columns the slot -> staged-column map can hold
constexpr int BB_MIN_QUERIES = 8;

__device__ __forceinline__ u32 bb_smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bb_mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bb_mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bb_mbar_arrive(u32 bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bb_mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "BB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra BB_DONE;\n"
      "bra BB_WAIT;\n"
      "BB_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bb_bulk_load(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bb_cp_async4(u32 dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bb_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct __align__(16) BbOp {  // one real token of a query's program (unknown tokens are dropped)
  double w;         // idf
  u32 arg;          // BB_OP_COL: byte offset of the staged column inside a stage; BB_OP_GLOBAL: dense slot;
                    // BB_OP_SPARSE: index of the token's BbCur / BbTile entry
  u32 op;
};
constexpr u32 BB_OP_COL = 0, BB_OP_GLOBAL = 1, BB_OP_SPARSE = 2;

struct __align__(16) BbCur {   // a sparse token's slice of the CURRENT tile (per query, in shared memory)
  u32 cur, end;                // unconsumed part: offsets into src
  const u32* src;              // in the query's arena, or in global memory when the slice did not fit
};
struct __align__(16) BbTile {  // a sparse token's tile-level state
  const u32* post;             // post_pack + term_ptr[t]; nullptr = not a sparse token
  const u32* skip;             // tile_skip row of the term
  u32 n0, n1;                  // slice bounds (offsets into post) of this CTA's next tile ...
  u32 m0, m1;                  // ... and of the one after (written by cp.async)
};

__device__ __forceinline__ double bb_below(double v) {  // largest double strictly below a finite v
  const long long bits = __double_as_longlong(v);
  if (v > 0.0) return __longlong_as_double(bits - 1);
  if (v < 0.0) return __longlong_as_double(bits + 1);
  return -4.9406564584124654e-324;
}

// Arena layout of one (query, tile): the slices [lo, hi) of the sparse tokens (lane = token)
// packed in token order; a slice that does not fit stays in global memory.  Returns the
// slice's offset in the arena or -1.
__device__ __forceinline__ int bb_arena_offset(u32 len, int lane) {
  u32 off = len;
#pragma unroll
  for (int d = 1; d < BB_MAXT; d <<= 1) {
    const u32 o = __shfl_up_sync(0xFFFFFFFFu, off, d);
    if (lane >= d) off += o;
  }
  off -= len;  // exclusive prefix
  return (len > 0 && off + len <= (u32)BB_ARENA) ? (int)off : -1;
}

__global__ void __launch_bounds__(BB_THREADS, 1)
bm25_batch_kernel(cmr_lex_index ix, const int* __restrict__ q_terms, const int* __restrict__ q_ptr,
                  const uint8_t* __restrict__ row_mask, KeyD* __restrict__ part, int n_queries, int chunk_q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_cols = reinterpret_cast<double*>(smem_raw);                      // [STAGES][NC][SUB]
  double* s_acc = s_cols + (size_t)BB_STAGES * BB_NC * BB_SUB;               // [CWARPS][SUB]
  KeyD* s_lists = reinterpret_cast<KeyD*>(s_acc + BB_CWARPS * BB_SUB);       // [QB][KP]
  BbOp* s_prog = reinterpret_cast<BbOp*>(s_lists + BB_QB * BB_KP);           // [QB][MAXT] token programs
  BbCur* s_cur = reinterpret_cast<BbCur*>(s_prog + BB_QB * BB_MAXT);         // [QB][MAXT]
  BbTile* s_tile = reinterpret_cast<BbTile*>(s_cur + BB_QB * BB_MAXT);       // [QB][MAXT]
  u32* s_arena = reinterpret_cast<u32*>(s_tile + BB_QB * BB_MAXT);           // [QB][2][ARENA]
  double* s_thr = reinterpret_cast<double*>(s_arena + (size_t)BB_QB * 2 * BB_ARENA);  // [QB] admission thresholds
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_thr + BB_QB);   // full[STAGES], empty[STAGES]
  int* s_colslot = reinterpret_cast<int*>(s_bar + 2 * BB_STAGES);            // [NC] dense slot of staged column c
  int* s_ncols = s_colslot + BB_NC;                                          // [1] (+3 pad)
  int* s_cost = s_ncols + 4;                                                 // [QB] estimated per-slice cost of a query
  int* s_order = s_cost + BB_QB;                                             // [QB] chunk-local query, heaviest first
  int* s_np = s_order + BB_QB;                                               // [QB] program length, -1 = not served here
  int* s_qid = s_np + BB_QB;                                                 // [QB] query of the slot
  short* s_colmap = reinterpret_cast<short*>(s_qid + BB_QB);                 // [n_dense] staged index, -1 unused, -2 not staged

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.y * chunk_q;
  const int q1 = (q0 + chunk_q) < n_queries ? (q0 + chunk_q) : n_queries;
  const u32 bar_full = bb_smem_u32(s_bar), bar_empty = bb_smem_u32(s_bar + BB_STAGES);
  const int G = (int)gridDim.x;

  // ---- set-up: lists, barriers, the chunk's distinct dense columns, query -> warp pairing ----
  for (int i = tid; i < BB_QB * BB_KP; i += BB_THREADS) key_clear(s_lists[i]);
  for (int i = tid; i < ix.n_dense; i += BB_THREADS) s_colmap[i] = -1;
  if (tid == 0) {
    for (int s = 0; s < BB_STAGES; ++s) {
      bb_mbar_init(bar_full + 8 * s, 1);
      bb_mbar_init(bar_empty + 8 * s, BB_CWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (ix.n_dense > 0) {
    const int t_lo = q_ptr[q0], t_hi = q_ptr[q1];
    for (int i = t_lo + tid; i < t_hi; i += BB_THREADS) {
      const int t = q_terms[i];
      if (t >= 0 && t < ix.n_terms) {
        const int slot = ix.dense_slot[t];
        if (slot >= 0) s_colmap[slot] = 0;  // benign race: every writer stores 0
      }
    }
  }
  if (tid < BB_QB) {
    // cost of one 256-document slice of this query: a dense token ~1 unit, a sparse token whose
    // list reaches most slices ~6 units (spill, scatter, reload), a rare one ~1
    int cost = -1;
    const int q = q0 + tid;
    if (q < q1) {
      cost = 0;
      const int lo = q_ptr[q], hi = q_ptr[q + 1];
      if (hi - lo <= BB_MAXT) {
        for (int i = lo; i < hi; ++i) {
          const int t = q_terms[i];
          if (t < 0 || t >= ix.n_terms) continue;
          if (ix.dense_slot != nullptr && ix.dense_slot[t] >= 0) { cost += 1; continue; }
          const long long df = ix.term_ptr[t + 1] - ix.term_ptr[t];
          cost += (df * BB_SUB >= ix.n_docs / 2) ? 6 : 1;
        }
      }
    }
    s_cost[tid] = cost;
  }
  __syncthreads();
  if (warp == 0) {
    int n = 0;
    for (int base = 0; base < ix.n_dense; base += 32) {
      const int i = base + lane;
      const bool used = i < ix.n_dense && s_colmap[i] == 0;
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, used);
      if (used) {
        const int idx = n + __popc(bal & ((1u << lane) - 1u));
        if (idx < BB_NC) {
          s_colmap[i] = (short)idx;
          s_colslot[idx] = i;
        } else {
          s_colmap[i] = -2;
        }
      }
      n += __popc(bal);
    }
    if (lane == 0) s_ncols[0] = n < BB_NC ? n : BB_NC;
  } else if (warp == 1) {
    const int mine = s_cost[lane];
    int rank = 0;
    for (int o = 0; o < BB_QB; ++o) {
      const int c = s_cost[o];
      rank += (c > mine) || (c == mine && o < lane);
    }
    s_order[rank] = lane;  // a permutation of 0..31; absent queries (cost -1) come last
  }
  __syncthreads();
  const int ncols = s_ncols[0];
  const int subs = ix.tile_docs / BB_SUB;

  if (warp == BB_CWARPS) {
    // ---- producer: stage the next slice of every column the chunk uses ---------------------
    u32 it = 0;
    for (int tile = blockIdx.x; tile < ix.n_tiles; tile += G) {
      const long long tile_lo = (long long)tile * ix.tile_docs;
      for (int s = 0; s < subs; ++s, ++it) {
        const long long doc0 = tile_lo + (long long)s * BB_SUB;
        if (doc0 >= ix.n_docs) break;
        const long long rem = ix.n_docs - doc0;
        const u32 n_here = rem < BB_SUB ? (u32)rem : (u32)BB_SUB;
        const u32 st = it % BB_STAGES, ph = (it / BB_STAGES) & 1u;
        bb_mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        if (lane == 0) bb_mbar_expect_tx(bar_full + 8 * st, (u32)ncols * n_here * 8u);
        __syncwarp();
        for (int c = lane; c < ncols; c += 32)
          bb_bulk_load(bb_smem_u32(s_cols + ((size_t)st * BB_NC + c) * BB_SUB),
                       ix.dense_imp + (size_t)s_colslot[c] * ix.n_docs + doc0, n_here * 8u, bar_full + 8 * st);
      }
    }
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------
  // Warp w serves the w-th heaviest and the w-th lightest query of the chunk: slots w and
  // w + CWARPS.  Everything a query carries from slice to slice lives in shared memory, so the
  // slice loop below needs few registers and its eight loads per dense token issue back to back.
  for (int qi = 0; qi < BB_QW; ++qi) {
    const int ls = warp + qi * BB_CWARPS;
    const int q = q0 + s_order[qi == 0 ? warp : BB_QB - 1 - warp];
    int np = -1;
    BbTile tt;
    tt.post = nullptr; tt.skip = nullptr; tt.n0 = tt.n1 = tt.m0 = tt.m1 = 0;
    if (q < q1) {
      const int qlo = q_ptr[q], m = q_ptr[q + 1] - qlo;
      if (m <= BB_MAXT) {
        int kind = 0, arg = 0;  // 0 nothing, 1 staged dense column, 2 dense column in global memory, 3 sparse
        double w = 0.0;
        if (lane < m) {
          const int t = q_terms[qlo + lane];
          if (t >= 0 && t < ix.n_terms) {
            w = ix.idf[t];
            const int slot = ix.dense_slot != nullptr ? ix.dense_slot[t] : -1;
            if (slot >= 0) {
              const int cm = s_colmap[slot];
              kind = cm >= 0 ? 1 : 2;
              arg = cm >= 0 ? cm : slot;
            } else {
              kind = 3;
              tt.post = ix.post_pack + ix.term_ptr[t];
              tt.skip = ix.tile_skip + (size_t)t * (ix.n_tiles + 1);
              const int first = blockIdx.x;  // < n_tiles (the grid is clamped)
              tt.n0 = tt.skip[first];
              tt.n1 = tt.skip[first + 1];
              if (first + G < ix.n_tiles) {
                tt.m0 = tt.skip[first + G];
                tt.m1 = tt.skip[first + G + 1];
              }
            }
          }
        }
        // the query's program: its real tokens in order, read back with one broadcast LDS.128 each
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, kind != 0);
        np = __popc(bal);
        if (kind != 0) {
          BbOp o;
          o.w = w;
          o.op = kind == 1 ? BB_OP_COL : (kind == 2 ? BB_OP_GLOBAL : BB_OP_SPARSE);
          o.arg = kind == 1 ? (u32)arg * (u32)(BB_SUB * 8) : (kind == 2 ? (u32)arg : (u32)lane);
          s_prog[ls * BB_MAXT + __popc(bal & ((1u << lane) - 1u))] = o;
        }
        // the first tile's postings
        const int off = bb_arena_offset(tt.post != nullptr ? tt.n1 - tt.n0 : 0u, lane);
        unsigned todo = __ballot_sync(0xFFFFFFFFu, off >= 0);
        u32* arena = s_arena + (size_t)(ls * 2 + 0) * BB_ARENA;
        while (todo) {
          const int j = __ffs(todo) - 1;
          todo &= todo - 1;
          const u32 l = __shfl_sync(0xFFFFFFFFu, tt.n1 - tt.n0, j);
          const u32* g = reinterpret_cast<const u32*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(tt.post + tt.n0), j));
          const u32 dst = bb_smem_u32(arena + __shfl_sync(0xFFFFFFFFu, off, j));
          for (u32 i = lane; i < l; i += 32) bb_cp_async4(dst + 4u * i, g + i);
        }
      }
    }
    if (lane < BB_MAXT) s_tile[ls * BB_MAXT + lane] = tt;
    if (lane == 0) {
      s_np[ls] = np;
      s_qid[ls] = q;
      s_thr[ls] = -INFINITY;
    }
  }
  __syncwarp();
  double* w_acc = s_acc + warp * BB_SUB;
  bool first_slice = true;

  u32 it = 0;
  u32 par = 0;  // arena buffer of the current tile
  for (int tile = blockIdx.x; tile < ix.n_tiles; tile += G, par ^= 1u) {
    const long long tile_lo = (long long)tile * ix.tile_docs;
    bb_cp_async_wait_all();
    __syncwarp();
    // Tile switch (lane = token): this tile's slices become current; the next tile's postings
    // and the skip entries of the tile after that are requested now and land during this tile.
    for (int qi = 0; qi < BB_QW; ++qi) {
      const int ls = warp + qi * BB_CWARPS;
      if (s_np[ls] < 0) continue;
      BbTile tt;
      tt.post = nullptr; tt.skip = nullptr; tt.n0 = tt.n1 = tt.m0 = tt.m1 = 0;
      if (lane < BB_MAXT) tt = s_tile[ls * BB_MAXT + lane];
      const bool sparse = tt.post != nullptr;
      {
        const u32 len = sparse ? tt.n1 - tt.n0 : 0u;
        const int off = bb_arena_offset(len, lane);  // where the staging pass put it
        if (lane < BB_MAXT) {
          BbCur c;
          c.cur = 0;
          c.end = len;
          c.src = off >= 0 ? s_arena + (size_t)(ls * 2 + par) * BB_ARENA + off : tt.post + tt.n0;
          s_cur[ls * BB_MAXT + lane] = c;
        }
      }
      tt.n0 = tt.m0;
      tt.n1 = tt.m1;
      if (tile + G < ix.n_tiles) {
        const u32 len = sparse ? tt.n1 - tt.n0 : 0u;
        const int off = bb_arena_offset(len, lane);
        unsigned todo = __ballot_sync(0xFFFFFFFFu, off >= 0);
        u32* arena = s_arena + (size_t)(ls * 2 + (par ^ 1u)) * BB_ARENA;
        while (todo) {
          const int j = __ffs(todo) - 1;
          todo &= todo - 1;
          const u32 l = __shfl_sync(0xFFFFFFFFu, len, j);
          const u32* g = reinterpret_cast<const u32*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(tt.post + tt.n0), j));
          const u32 dst = bb_smem_u32(arena + __shfl_sync(0xFFFFFFFFu, off, j));
          for (u32 i = lane; i < l; i += 32) bb_cp_async4(dst + 4u * i, g + i);
        }
      }
      if (sparse) {
        BbTile* dstt = s_tile + ls * BB_MAXT + lane;
        dstt->n0 = tt.n0;
        dstt->n1 = tt.n1;
        if (tile + 2 * G < ix.n_tiles) {
          bb_cp_async4(bb_smem_u32(&dstt->m0), tt.skip + tile + 2 * G);
          bb_cp_async4(bb_smem_u32(&dstt->m1), tt.skip + tile + 2 * G + 1);
        }
      }
    }
    __syncwarp();
    for (int s = 0; s < subs; ++s, ++it) {
      const long long doc0 = tile_lo + (long long)s * BB_SUB;
      if (doc0 >= ix.n_docs) break;
      const long long rem = ix.n_docs - doc0;
      const int n_here = rem < BB_SUB ? (int)rem : BB_SUB;
      const u32 st = it % BB_STAGES, ph = (it / BB_STAGES) & 1u;
      const u32 loc_lo = (u32)(s * BB_SUB), loc_hi = loc_lo + BB_SUB;
      bb_mbar_wait(bar_full + 8 * st, ph);
      const unsigned char* cols_b = reinterpret_cast<const unsigned char*>(s_cols + (size_t)st * BB_NC * BB_SUB) + lane * 8;
      for (int qi = 0; qi < BB_QW; ++qi) {
        const int ls = warp + qi * BB_CWARPS;
        const int np = s_np[ls];
        if (np < 0) continue;  // warp-uniform
        double a[BB_R];
#pragma unroll
        for (int r = 0; r < BB_R; ++r) a[r] = 0.0;
        bool spilled = false;  // the warp's strip in shared memory, not a[], holds the accumulators
        const BbOp* prog = s_prog + ls * BB_MAXT;
        for (int t = 0; t < np; ++t) {
          const BbOp o = prog[t];  // broadcast LDS.128
          const double w = o.w;
          if (o.op != BB_OP_SPARSE) {
            if (spilled) {
              __syncwarp();
#pragma unroll
              for (int r = 0; r < BB_R; ++r) a[r] = w_acc[r * 32 + lane];
              spilled = false;
            }
            double f[BB_R];
            if (o.op == BB_OP_COL) {
              const double* cp = reinterpret_cast<const double*>(cols_b + o.arg);
#pragma unroll
              for (int r = 0; r < BB_R; ++r) f[r] = cp[r * 32];
            } else {
              const double* gp = ix.dense_imp + (size_t)o.arg * ix.n_docs + doc0 + lane;
#pragma unroll
              for (int r = 0; r < BB_R; ++r) f[r] = (r * 32 + lane < n_here) ? ldg_stream_f64(gp + r * 32) : 0.0;
            }
#pragma unroll
            for (int r = 0; r < BB_R; ++r) a[r] = __dadd_rn(a[r], __dmul_rn(w, f[r]));  // no fma (rank_bm25 rounds the product)
            continue;
          }
          // sparse token: anything of its slice inside these 256 documents?
          BbCur* cs = s_cur + ls * BB_MAXT + o.arg;
          const BbCur cc = *cs;  // broadcast LDS.128
          u32 c = cc.cur;
          const u32 e = cc.end;
          if (c >= e) continue;
          u32 pk = c + lane < e ? cc.src[c + lane] : 0xFFFFFFFFu;
          if ((__shfl_sync(0xFFFFFFFFu, pk, 0) & 0xFFFFu) >= loc_hi) continue;  // sorted by document
          if (!spilled) {
#pragma unroll
            for (int r = 0; r < BB_R; ++r) w_acc[r * 32 + lane] = a[r];
            spilled = true;
          }
          __syncwarp();  // the strip is complete (spill or the previous token's scatter)
          for (;;) {
            const bool in = c + lane < e && (pk & 0xFFFFu) < loc_hi;  // a prefix of the lanes
            if (in) {
              const double imp = __ldg(ix.imp_table + (pk >> 16));
              double* p = w_acc + ((pk & 0xFFFFu) - loc_lo);
              *p = __dadd_rn(*p, __dmul_rn(w, imp));  // documents are unique within a list
            }
            const int cnt = __popc(__ballot_sync(0xFFFFFFFFu, in));
            c += (u32)cnt;
            if (cnt < 32) break;
            pk = c + lane < e ? cc.src[c + lane] : 0xFFFFFFFFu;
          }
          if (lane == 0) cs->cur = c;  // read again at the next slice (a __syncwarp away)
        }
        if (spilled) {
          __syncwarp();
#pragma unroll
          for (int r = 0; r < BB_R; ++r) a[r] = w_acc[r * 32 + lane];
          __syncwarp();  // the strip may be rewritten by the next query
        }

        // ---- selection straight from the registers ------------------------------------------
        KeyD* list = s_lists + (size_t)ls * BB_KP;
        double thr = s_thr[ls];
        const double thr_in = thr;
        if (first_slice) {
          // first slice of this CTA: each lane's best score is reached by a distinct document,
          // so the smallest of the 32 lane maxima is a lower bound of the CTA's 32nd best score
          double mx = -INFINITY;
#pragma unroll
          for (int r = 0; r < BB_R; ++r) {
            const int i = r * 32 + lane;
            bool ok = i < n_here;
            if (ok && row_mask != nullptr) ok = row_mask[doc0 + i] != 0;
            if (ok) mx = fmax(mx, a[r]);
          }
          double mn = mx;
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, off));
          if (mn > -INFINITY) thr = bb_below(mn);
        }
        {
          // Fast reject: for a threshold >= +0.0, x > thr needs x > 0 and then the bit patterns
          // compare like signed integers, so "no high word reaches the threshold's" proves no hit.
          // (A negative threshold -- the list is not full yet -- takes the exact path below.)
          const int t_hi = __double2hiint(thr);
          int mx = __double2hiint(a[0]);
#pragma unroll
          for (int r = 1; r < BB_R; ++r) mx = max(mx, __double2hiint(a[r]));
          const bool any_hit = t_hi < 0 || mx >= t_hi;
          if (__any_sync(0xFFFFFFFFu, any_hit)) {
#pragma unroll
            for (int r = 0; r < BB_R; ++r) {
              const int i = r * 32 + lane;
              bool ok = (i < n_here) && (a[r] > thr);
              if (ok && row_mask != nullptr) ok = row_mask[doc0 + i] != 0;
              unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
              while (bal) {
                const int src_lane = __ffs(bal) - 1;
                bal &= bal - 1;
                KeyD key;
                key.s = __shfl_sync(0xFFFFFFFFu, a[r], src_lane);
                key.id = (u32)(doc0 + r * 32 + src_lane);
                key.pad = 0;
                if (key.s > thr) {
                  KeyD new_last;
                  key_clear(new_last);
                  if (warp_list_insert<BB_KP, KeyD>(list, key, lane, new_last) && !key_empty(new_last)) thr = new_last.s;
                }
              }
            }
          }
        }
        if (lane == 0 && thr != thr_in) s_thr[ls] = thr;
      }
      first_slice = false;
      __syncwarp();
      if (lane == 0) bb_mbar_arrive(bar_empty + 8 * st);
    }
  }

  for (int qi = 0; qi < BB_QW; ++qi) {
    const int ls = warp + qi * BB_CWARPS;
    if (s_np[ls] < 0) continue;
    KeyD* dst = part + ((size_t)s_qid[ls] * gridDim.x + blockIdx.x) * BB_KP;
    dst[lane] = s_lists[(size_t)ls * BB_KP + lane];
  }
}

static inline size_t bb_smem_bytes(int n_dense) {
  return (size_t)BB_STAGES * BB_NC * BB_SUB * 8 + (size_t)BB_CWARPS * BB_SUB * 8 + (size_t)BB_QB * BB_KP * sizeof(KeyD) +
         (size_t)BB_QB * BB_MAXT * (sizeof(BbOp) + sizeof(BbCur) + sizeof(BbTile)) + (size_t)BB_QB * 2 * BB_ARENA * 4 +
         (size_t)BB_QB * 8 + 2 * BB_STAGES * 8 + BB_NC * 4 + 16 + 4 * BB_QB * 4 + (size_t)((n_dense + 7) / 8 * 8) * 2 + 128;
}

struct Bm25Plan {
  bool packed;
  tile_fn_t fn;
  int kpl;
  int grid_y;  // tile groups (lists per query)
  size_t smem_tile, smem_fin;
  bool batch;          // bm25_batch_kernel serves the queries of up to BB_MAXT tokens
  int bb_chunks, bb_chunk_q;
  size_t smem_bb;
};

static bool batch_enabled() {
  const char* e = getenv("CMR_BM25_BATCH");  // read per call: cheap, and a process may change it
  return e && e[0] == '1';                   // opt-in: see the note above bm25_batch_kernel
}

static int check_index(const cmr_lex_index* ix) {
  CMR_CHECK_ARG(ix != nullptr, "null index");
  CMR_CHECK_ARG(ix->n_docs >= 0 && ix->n_docs < 0xFFFFFFFFll, "n_docs out of range");
  CMR_CHECK_ARG(ix->tile_docs >= 512 && ix->tile_docs % 512 == 0 && ix->tile_docs <= 65536,
                "tile_docs must be a multiple of 512 in [512, 65536]");
  CMR_CHECK_ARG(ix->n_terms == 0 || (ix->post_pack && ix->imp_table) || (ix->post_doc && ix->post_imp) || ix->term_ptr,
                "index needs packed or wide postings");
  CMR_CHECK_ARG(ix->n_tiles >= 1 && (long long)ix->n_tiles * ix->tile_docs >= ix->n_docs, "n_tiles inconsistent with n_docs/tile_docs");
  CMR_CHECK_ARG(ix->n_terms >= 0, "n_terms negative");
  CMR_CHECK_ARG(ix->n_terms == 0 || (ix->term_ptr && ix->tile_skip && ix->idf), "null index arrays");
  CMR_CHECK_ARG(ix->n_dense >= 0 && (ix->n_dense == 0 || (ix->dense_imp && ix->dense_slot)), "dense columns inconsistent");
  return CMR_OK;
}

static int make_plan(const cmr_lex_index& ix, int n_queries, int k, Bm25Plan* p) {
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
  // batched form: one persistent CTA per SM, the lists of a query are the CTAs
  p->batch = batch_enabled() && n_queries >= BB_MIN_QUERIES && p->kpl == 1 && p->packed && ix.n_terms > 0 &&
             ix.n_dense <= BB_MAX_DENSE && ix.tile_docs % BB_SUB == 0 &&
             (ix.n_dense == 0 || (ix.n_docs % 2 == 0 && ((uintptr_t)ix.dense_imp % 16) == 0));
  p->bb_chunks = p->bb_chunk_q = 0;
  p->smem_bb = 0;
  if (p->batch) {
    p->bb_chunks = (n_queries + BB_QB - 1) / BB_QB;
    p->bb_chunk_q = (n_queries + p->bb_chunks - 1) / p->bb_chunks;
    p->smem_bb = bb_smem_bytes(ix.n_dense);
    p->grid_y = sms < ix.n_tiles ? sms : ix.n_tiles;
    if (p->bb_chunks > 65535) p->batch = false;
  }
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
  int min_tokens = -1;
  if (p.batch) {
    static int bb_attr_mask = 0;
    if (!(bb_attr_mask & (1 << dev))) {
      cudaError_t e = cudaFuncSetAttribute(bm25_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(bm25_batch)");
      bb_attr_mask |= (1 << dev);
    }
    bm25_batch_kernel<<<dim3(p.grid_y, p.bb_chunks), BB_THREADS, p.smem_bb, st>>>(ix, q_terms, q_ptr, row_mask, part,
                                                                                 n_queries, p.bb_chunk_q);
    min_tokens = BB_MAXT;  // the tile kernel only serves longer queries (its CTAs leave at once otherwise)
  }
  p.fn<<<grid, BM_THREADS, p.smem_tile, st>>>(ix, q_terms, q_ptr, row_mask, part, min_tokens);
  bm25_finalize_kernel<KPL><<<n_queries, BMF_THREADS, p.smem_fin, st>>>(part, p.grid_y, row_offset, k, out_scores,
                                                                       out_ids, out_counts, out_flags);
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
  if (kpl == 1) bm25_finalize_kernel<1><<<n_queries, BMF_THREADS, smem, st>>>(part, n_lists, row_offset, k, out_scores, out_ids, out_counts, out_flags);
  else if (kpl == 2) bm25_finalize_kernel<2><<<n_queries, BMF_THREADS, smem, st>>>(part, n_lists, row_offset, k, out_scores, out_ids, out_counts, out_flags);
  else bm25_finalize_kernel<4><<<n_queries, BMF_THREADS, smem, st>>>(part, n_lists, row_offset, k, out_scores, out_ids, out_counts, out_flags);
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
