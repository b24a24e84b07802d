// bm25_mma.cu -- A2 batched: BM25 top-k for a batch of queries with the head terms on the tensor cores.
//
// Replaces BM25Okapi.get_scores + sorted(...)[:k] inside BM25Store.search (reference
// rag/retrieval/bm25.py:175-212) for batches; the exact kernels of bm25.cu cost ~43 us per query
// at 10M documents because every query sweeps every document on its own.  Here the part of the
// score that every document has -- the contributions of the up to 64 terms with the longest
// posting lists ("head" terms: 97 % of the posting volume a Zipf query touches) -- is one matrix
// product per block of 32 queries:
//
//     approx[d, q] = sum_c head_mat[d, c] * count[q, c]  +  sum over sparse tokens t of q: fp16(idf_t * f(t, d))
//
// head_mat [n_docs, 64] holds fp16(idf * factor) (built with the index), count[q, c] is how often
// head term c occurs in query q (exact in fp16), the products are exact and the tensor pipe
// accumulates in fp32.  The sparse remainder is a CSR posting-list scatter:
//
//   bm25x_prep_kernel     per block of 32 queries: the count matrix and the list of (query,
//                         sparse term, idf) pairs; queries with more than 16 tokens or a negative
//                         idf are flagged (the exact kernels serve them).
//   bm25x_bucket_kernel   one CTA per tile of the index: the slices of the pairs' posting lists
//                         that fall into the tile (skip table, no search) are scattered into
//                         one bucket per 32 documents as (document, query, fp16 contribution)
//                         words; slots come from shared-memory counters, the words go straight
//                         to the bucket rows in HBM (~200 B per 32 documents at 32 queries: 5 %
//                         of the head matrix).
//   bm25x_mma_kernel      persistent, one CTA per SM.  warp 0: TMA producer, one [128 documents x
//                         64 terms] box of head_mat per item into an 8-stage ring of 128-byte
//                         swizzled tiles (16 KB each).  warp 1: one thread issues 4 x tcgen05.mma
//                         (cta_group::1, kind::f16, M = 128 documents = TMEM lanes, N = 32 queries =
//                         TMEM columns, K = 16) per item into a ring of 8 TMEM accumulators;
//                         tcgen05.commit frees the ring slot and publishes the accumulator.
//                         warps 2-9: epilogue, two warps per TMEM lane quarter on alternating items.
//                         A warp scatters its bucket (prefetched one item ahead) into a 32 x 32
//                         fp32 tile in shared memory, reads the accumulator with tcgen05.ld (one
//                         document per thread, 32 queries per load), adds its row of the tile and
//                         compares with the 32 admission bounds held in registers; the rare
//                         survivors are appended to the CTA's candidate list of the query.
//                         SAMPLE mode visits every 16th tile and keeps per-thread running maxima;
//                         the KP-th largest of those group maxima (bm25x_bound_kernel) is
//                         attained by KP different documents, hence a lower bound of the KP-th best
//                         approximate score: MAIN mode with that bound keeps a superset of the
//                         approximate top KP.
//   bm25x_finalize_kernel one CTA per query: the KP best candidates by approximate score are
//                         rescored exactly -- float64, query-token order, idf * factor rounded then
//                         added (rank_bm25's operation order), the factor of every (token, document)
//                         found by bisection in the term's tile slice -- and ranked (score desc,
//                         document asc).  Certificate: every document outside the KP has approximate
//                         score <= cut-off, hence exact score <= (cut-off + abs) / (1 - rho) with
//                         rho = 2^-11 (fp16 rounding of each contribution) + accumulation slack;
//                         if the k-th exact score is above that, the top k is THE top k and its
//                         scores are bit-identical to the exact kernels'.  Otherwise the query is
//                         flagged and bm25.cu re-runs it.
//
// Algorithmic bytes per block of 32 queries: n_docs * 128 (head_mat, read once) + 4 per sparse
// posting (read by the bucket kernel) + 2 * 4 per bucket entry (written, read back), + 1/16 of
// the matrix for the sample pass.  HBM-bound: the MMA work per item is 128 x 32 x 64.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "bm25_head.cuh"
#include "lex_slice.cuh"
#include "mma_common.cuh"

namespace cmr {

constexpr int BX_M = 128;                     // documents per item (UMMA M, TMEM lanes)
constexpr int BX_N = 32;                      // queries per block (UMMA N, TMEM columns)
constexpr int BX_K = 64;                      // head terms: one 128-byte swizzle span of fp16
#ifndef CMR_BX_STAGES
#define CMR_BX_STAGES 8
#endif
constexpr int BX_STAGES = CMR_BX_STAGES;      // ring of head_mat boxes (128 KB in flight per SM)
constexpr int BX_ACC = 16;                    // barrier slots for the TMEM accumulators (512 / (32 * query blocks) are in use)
constexpr int BX_A_BYTES = BX_M * BX_K * 2;   // 16 KiB
constexpr int BX_Q_BYTES = BX_N * BX_K * 2;   // 4 KiB
constexpr int BX_MAX_QB = 4;                  // blocks of 32 queries served by one pass over head_mat
#ifndef CMR_BX_EPI_GROUPS
#define CMR_BX_EPI_GROUPS 3
#endif
constexpr int BX_EPI_GROUPS = CMR_BX_EPI_GROUPS;   // epilogue warp groups (4 warps each) on alternating items
constexpr int BX_EPI_WARPS = 4 * BX_EPI_GROUPS;
constexpr int BX_THREADS = (2 + BX_EPI_WARPS) * 32;
constexpr int BX_S_BYTES = 32 * BX_N * 4;     // one epilogue warp's 32 documents x 32 queries fp32 tile
constexpr int BX_CAP = 256;                   // words per bucket of 32 documents; word 0 = entry count
constexpr int BX_MAXT = 16;                   // tokens per query served here
constexpr int BX_MAX_PAIRS = BX_N * BX_MAXT;  // (query, sparse term) pairs per block
constexpr int BX_SAMPLE = 0, BX_MAIN = 1;
constexpr int BX_SAMPLE_STRIDE = 16;
constexpr int BX_GROUPS_PER_CTA = 4;                  // SAMPLE: group maxima each CTA hands to the bound kernel (one per TMEM lane quarter)
constexpr int BX_LIST_CAP = 256;              // candidate slots per (CTA, query)
constexpr int BX_CAP_PER_KP = 128;            // candidates finalize can collect per query = 128 * KP
constexpr int BX_MAX_TILE_DOCS = 2048;        // bucket kernel: tile_docs / 32 buckets of BX_CAP words in shared memory
// kind::f16 instruction descriptor (built in the kernel: N depends on the query blocks of the pass): D = f32
// (bit 4), A = B = fp16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24

// reasons a query is handed back to the exact kernels (bit 0 = CMR_FLAG_UNCERTIFIED is always set with them)
constexpr int BX_F_LONG = 2, BX_F_NEG = 4, BX_F_BUCKET = 8, BX_F_LIST = 16, BX_F_CERT = 32;

// |approx - exact| <= BX_RHO * exact + BX_ABS for non-negative contributions: every contribution is
// rounded to fp16 once (2^-11 relative; below the fp16 normal range 2^-25 absolute), products with
// the integer counts are exact, at most 64 + 16 fp32 additions (2^-23 relative each, the tensor
// pipe may truncate) -> 2^-11 + 80 * 2^-23 < 2^-11 * 1.03; 1.0625 leaves a margin.  The sparse
// contributions are summed in 2^-16 fixed point: exact for fp16 values >= 2^-6, up to 2^-17 of
// rounding each below that (at most BX_MAXT of them), which BX_ABS covers.
constexpr double BX_RHO = 1.0625 / 2048.0;
constexpr double BX_ABS = 1.3e-4;

struct __align__(16) BxPair {
  int term;
  int q;      // query within the block
  double w;   // idf
};

__device__ __forceinline__ u32 ldg_stream_word(const u32* p) {
  u32 r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------
// One CTA per block of 32 queries, 16 lanes per query: a lane owns one token of a round of 16, so
// the term / idf / head-slot loads of a query are three dependent round trips, not 3 per token.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BX_N * 16)
bm25x_prep_kernel(cmr_lex_index ix, const int* __restrict__ q_terms, const int* __restrict__ q_ptr, int n_queries,
                  __half* __restrict__ qmat, BxPair* __restrict__ pairs, int* __restrict__ n_pairs,
                  int* __restrict__ flags) {
  __shared__ int s_cnt[BX_N][BX_K];
  __shared__ int s_np;
  const int blk = blockIdx.x, tid = threadIdx.x, q0 = blk * BX_N;
  const int ql = tid >> 4, sub = tid & 15;
  const unsigned half_mask = 0xFFFFu << (tid & 16);   // the 16 lanes of this query
  for (int i = tid; i < BX_N * BX_K; i += BX_N * 16) (&s_cnt[0][0])[i] = 0;
  if (tid == 0) s_np = 0;
  __syncthreads();
  if (q0 + ql < n_queries) {
    const int q = q0 + ql;
    const int lo = q_ptr[q], hi = q_ptr[q + 1];
    int flag = 0, real = 0;
    for (int base = lo; base < hi; base += 16) {
      const int i = base + sub;
      int t = -1;
      if (i < hi) t = q_terms[i];
      const bool valid = t >= 0 && t < ix.n_terms;
      double w = 0.0;
      int slot = -1;
      if (valid) {
        w = ix.idf[t];
        slot = ix.head_slot[t];
      }
      const unsigned vm = __ballot_sync(half_mask, valid) & half_mask;
      const int before = real + __popc(vm & ((1u << (tid & 31)) - 1u));   // valid tokens ahead of this one
      real += __popc(vm);
      if (valid && before < BX_MAXT) {
        if (w < 0.0) flag |= BX_F_NEG;
        if (w != 0.0) {
          if (slot >= 0) {
            atomicAdd(&s_cnt[ql][slot], 1);
          } else {
            const int p = atomicAdd(&s_np, 1);   // <= BX_MAXT per query
            BxPair pr;
            pr.term = t;
            pr.q = ql;
            pr.w = w;
            pairs[(size_t)blk * BX_MAX_PAIRS + p] = pr;
          }
        }
      }
    }
    if (real > BX_MAXT) flag |= BX_F_LONG;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) flag |= __shfl_xor_sync(half_mask, flag, off);
    if (sub == 0) flags[q] = flag ? (CMR_FLAG_UNCERTIFIED | flag) : 0;
  }
  __syncthreads();
  for (int i = tid; i < BX_N * BX_K; i += BX_N * 16) qmat[(size_t)blk * BX_N * BX_K + i] = __int2half_rn((&s_cnt[0][0])[i]);
  if (tid == 0) n_pairs[blk] = s_np;
}

// ---------------------------------------------------------------------------------------
// One CTA per tile of the index: bucket the block's sparse postings by 32-document group.
// buckets [n_tiles * tile_docs / 32][BX_CAP]: word 0 = number of entries (<= BX_CAP - 1), then
// entries  bits 0-4 document within the group | bits 5-9 query | bits 16-31 fp16(idf * factor).
// Slots are handed out by shared-memory counters; the entries go straight to the bucket rows in
// global memory (all writers of a row are this CTA, within microseconds: L2 merges the sectors).
// ---------------------------------------------------------------------------------------
#ifndef CMR_BXB_U
#define CMR_BXB_U 4
#endif
#ifndef CMR_BXB_THREADS
#define CMR_BXB_THREADS 256
#endif
constexpr int BXB_THREADS = CMR_BXB_THREADS;
constexpr int BXB_U = CMR_BXB_U;   // postings in flight per thread

__global__ void __launch_bounds__(BXB_THREADS)
bm25x_bucket_kernel(cmr_lex_index ix, const BxPair* __restrict__ pairs_all, const int* __restrict__ n_pairs_all,
                    int q_base0, u32* __restrict__ buckets_all, size_t bucket_stride, int* __restrict__ flags) {
  // grid (index tile, query block of the group)
  const BxPair* pairs = pairs_all + (size_t)blockIdx.y * BX_MAX_PAIRS;
  const int* n_pairs_p = n_pairs_all + blockIdx.y;
  const int q_base = q_base0 + (int)blockIdx.y * BX_N;
  u32* buckets = buckets_all + (size_t)blockIdx.y * bucket_stride;
  // the pairs that have postings in this tile, compacted: slice start, running end offset, idf, query
  __shared__ long long s_lo[BX_MAX_PAIRS];
  __shared__ int s_off[BX_MAX_PAIRS + 1];
  __shared__ double s_w[BX_MAX_PAIRS];
  __shared__ int s_q[BX_MAX_PAIRS];
  __shared__ int s_cnt[BX_MAX_TILE_DOCS / 32];
  __shared__ int s_warp_sum[BXB_THREADS / 32], s_warp_live[BXB_THREADS / 32];
  __shared__ int s_next, s_np;
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_b = ix.tile_docs / 32;
  const int np = *n_pairs_p;
  u32* rows = buckets + (size_t)tile * n_b * BX_CAP;
  for (int b = tid; b < n_b; b += BXB_THREADS) s_cnt[b] = 0;
  if (tid == 0) {
    s_off[0] = 0;
    s_next = 0;
  }
  // Slices of up to BXB_THREADS pairs per round: every thread looks one up (two dependent loads),
  // then a block scan hands the non-empty ones their place in the compact list.
  int carry = 0, live = 0;   // postings / non-empty pairs of the rounds so far (uniform)
  for (int p0 = 0; p0 < np; p0 += BXB_THREADS) {
    const int p = p0 + tid;
    u32 a = 0, z = 0;
    BxPair pr;
    pr.term = 0;
    pr.q = 0;
    pr.w = 0.0;
    if (p < np) {
      pr = pairs[p];
      lex_slice(ix, pr.term, lex_skip_row(ix, pr.term), tile, &a, &z);
    }
    const int len = (int)(z - a);
    int incl = len;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off) incl += o;
    }
    const unsigned nz = __ballot_sync(0xFFFFFFFFu, len > 0);
    if (lane == 31) s_warp_sum[warp] = incl;
    if (lane == 0) s_warp_live[warp] = __popc(nz);
    __syncthreads();
    int sum_before = carry, live_before = live, sum_all = carry, live_all = live;
#pragma unroll
    for (int w = 0; w < BXB_THREADS / 32; ++w) {
      if (w < warp) {
        sum_before += s_warp_sum[w];
        live_before += s_warp_live[w];
      }
      sum_all += s_warp_sum[w];
      live_all += s_warp_live[w];
    }
    if (len > 0) {
      const int ci = live_before + __popc(nz & ((1u << lane) - 1u));
      s_lo[ci] = ix.term_ptr[pr.term] + a;
      s_off[ci + 1] = sum_before + incl;
      s_w[ci] = pr.w;
      s_q[ci] = pr.q;
    }
    carry = sum_all;
    live = live_all;
    __syncthreads();   // s_warp_* are rewritten by the next round
  }
  if (tid == 0) s_np = live;
  __syncthreads();
  const int npc = s_np;
  const int total = s_off[npc];
  // A warp takes chunks of 32 * BXB_U consecutive postings of the concatenated slices (handed out
  // by a counter, so no warp idles while another has a chunk left): one bisection for the chunk's
  // first posting, then every lane walks forward over the (non-empty) slices from there.
  constexpr int CH = 32 * BXB_U;
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_next, 1) * CH;
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= total) break;
    int p = 0;
    {
      int hi = npc;   // the pair whose slice holds posting `base`: largest p with s_off[p] <= base
      while (hi - p > 1) {
        const int mid = (p + hi) >> 1;
        if (s_off[mid] <= base) p = mid;
        else hi = mid;
      }
    }
    int pr_of[BXB_U];
    u32 pk[BXB_U];
    double imp[BXB_U];
#pragma unroll
    for (int u = 0; u < BXB_U; ++u) {
      const int i = base + u * 32 + lane;
      pk[u] = 0;
      if (i < total) {
        while (s_off[p + 1] <= i) ++p;
        pk[u] = ldg_stream_word(ix.post_pack + s_lo[p] + (i - s_off[p]));
      }
      pr_of[u] = p;
    }
#pragma unroll
    for (int u = 0; u < BXB_U; ++u) imp[u] = __ldg(ix.imp_table + (pk[u] >> 16));   // all gathers in flight together
#pragma unroll
    for (int u = 0; u < BXB_U; ++u) {
      if (base + u * 32 + lane >= total) break;
      const u32 local = pk[u] & 0xFFFFu;
      const int q = s_q[pr_of[u]];
      const __half h = __double2half(__dmul_rn(s_w[pr_of[u]], imp[u]));
      const int b = (int)(local >> 5);
      const int slot = atomicAdd(&s_cnt[b], 1) + 1;
      if (slot < BX_CAP) rows[b * BX_CAP + slot] = (local & 31u) | ((u32)q << 5) | ((u32)__half_as_ushort(h) << 16);
      else atomicOr(&flags[q_base + q], CMR_FLAG_UNCERTIFIED | BX_F_BUCKET);
    }
  }
  __syncthreads();
  for (int b = tid; b < n_b; b += BXB_THREADS) rows[b * BX_CAP] = (u32)(s_cnt[b] < BX_CAP - 1 ? s_cnt[b] : BX_CAP - 1);
}

// ---------------------------------------------------------------------------------------
struct BxParams {
  long long n_docs;
  int n_items;          // tiles of 128 documents this launch visits
  int stride;           // SAMPLE: visited tile = item * stride (full tiles only); MAIN: 1
  int n_qb;             // blocks of 32 queries in this group (MAIN: all in one pass; SAMPLE: blockIdx.y picks one)
  int n_queries;        // queries of the group (<= 32 * n_qb)
  const u32* buckets;   // [n_qb][bucket_stride words]: per block [ceil(n_docs / 32) up to whole index tiles][BX_CAP]
  size_t bucket_stride;
  float* gmax;          // SAMPLE: [32 * n_qb][gridDim.x * BX_GROUPS_PER_CTA] maxima of every group, per query
  const float* thr;     // MAIN: [32 * n_qb] admission bounds
  u64* cand;            // MAIN: [gridDim.x][32 * n_qb][cap] keys (orderable fp32 score, ~document)
  int* cnt;             // MAIN: [gridDim.x][32 * n_qb] entries appended (may exceed cap)
  int cap;
};

// shared memory (offsets from the 1024-byte aligned base): ring, query tiles, epilogue tiles, control
constexpr u32 BX_OFF_Q = BX_STAGES * BX_A_BYTES;
constexpr u32 BX_OFF_S = BX_OFF_Q + BX_MAX_QB * BX_Q_BYTES;
constexpr u32 BX_OFF_BAR = BX_OFF_S + BX_EPI_WARPS * BX_S_BYTES;
// barriers: full[s] +8s, empty[s] +64+8s, tfull[a] +128+8a, tempty[a] +256+8a, qfull +384; TMEM base +392
constexpr u32 BX_BAR_TFULL = 128, BX_BAR_TEMPTY = 256, BX_BAR_QFULL = 384, BX_BAR_TMEM = 392;
static_assert(BX_STAGES <= 8 && BX_ACC <= 16, "barrier slots");
constexpr u32 BX_OFF_THR = BX_OFF_BAR + 400;                       // [32 * BX_MAX_QB] admission bounds
constexpr u32 BX_OFF_CNT = BX_OFF_THR + BX_MAX_QB * BX_N * 4;      // [32 * BX_MAX_QB] list counters
constexpr size_t BX_SMEM = 1024 + BX_OFF_CNT + BX_MAX_QB * BX_N * 4;

// Work order shared by the three roles.  A CTA owns items first, first + step, ...; item number li
// (0, 1, ...) gets ONE accumulator of 32 * nqb TMEM columns (accumulator li % n_acc, phase (li / n_acc) & 1,
// n_acc = 512 / (32 * nqb)): the MMAs of an item cover all its query blocks at once (UMMA N = 32 * nqb).
// A tcgen05.mma of this shape costs ~150 cycles of latency whatever its N and the 4 MMAs of a K = 64 chain
// accumulate into the same columns, i.e. serially -- with one chain per (item, block) the issuing thread, not
// HBM or the epilogue, bounded the kernel (a 128-column head matrix doubled its time).  Epilogue group g
// serves the items with li % BX_EPI_GROUPS == g, block by block (use = (item, block): the block's 32 columns).
// A parity wait must never be a whole phase behind: a warp that waits for item li has seen the commit of its
// previous item li - BX_EPI_GROUPS, and the commit of li - n_acc (the item before li on the same accumulator)
// precedes that one only if BX_EPI_GROUPS <= n_acc; n_acc >= 4, hence the assert.
static_assert(BX_EPI_GROUPS <= 512 / (BX_N * BX_MAX_QB), "accumulator ring too shallow for the epilogue groups");
template <int MODE>
__global__ void __launch_bounds__(BX_THREADS, 1)
bm25x_mma_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_rows,
                 const BxParams p) {
  extern __shared__ unsigned char smem_raw[];
  const u32 raw = smem_u32(smem_raw);
  const u32 base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gen = smem_raw + (base - raw);
  const u32 bars = base + BX_OFF_BAR;
  volatile u32* tmem_slot = reinterpret_cast<volatile u32*>(gen + BX_OFF_BAR + BX_BAR_TMEM);
  float* s_thr = reinterpret_cast<float*>(gen + BX_OFF_THR);
  int* s_cnt = reinterpret_cast<int*>(gen + BX_OFF_CNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // SAMPLE: one query block per CTA row of the grid; MAIN: every block of the group, head_mat read once
  const int nqb = MODE == BX_SAMPLE ? 1 : p.n_qb;
  const int qb0 = MODE == BX_SAMPLE ? (int)blockIdx.y : 0;

  for (int i = threadIdx.x; i < BX_EPI_WARPS * BX_S_BYTES / 16; i += BX_THREADS)
    reinterpret_cast<float4*>(gen + BX_OFF_S)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < BX_MAX_QB * BX_N; i += BX_THREADS) {
    s_cnt[i] = 0;
    s_thr[i] = (MODE == BX_MAIN && i < p.n_queries) ? p.thr[i] : INFINITY;
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_rows) : "memory");
    for (int s = 0; s < BX_STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 64 + 8 * s, 1);
    }
    for (int a = 0; a < BX_ACC; ++a) {
      mbar_init(bars + BX_BAR_TFULL + 8 * a, 1);
      mbar_init(bars + BX_BAR_TEMPTY + 8 * a, 4);  // one arrival per epilogue warp of the group that drains the item
    }
    mbar_init(bars + BX_BAR_QFULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the allocating warp also frees
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + BX_BAR_TMEM),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;
  const int first = (int)blockIdx.x, step = (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_expect_tx(bars + BX_BAR_QFULL, (u32)(nqb * BX_Q_BYTES));
      for (int qb = 0; qb < nqb; ++qb)
        tma_load_2d(base + BX_OFF_Q + qb * BX_Q_BYTES, &tm_q, bars + BX_BAR_QFULL, 0, (qb0 + qb) * BX_N, TMA_EVICT_LAST);
      u32 s = 0, ph = 0;
      for (int it = first; it < p.n_items; it += step) {
        mbar_wait(bars + 64 + 8 * s, ph ^ 1u);
        mbar_expect_tx(bars + 8 * s, BX_A_BYTES);
        tma_load_2d(base + s * BX_A_BYTES, &tm_rows, bars + 8 * s, 0, it * p.stride * BX_M, TMA_EVICT_FIRST);
        if (++s == BX_STAGES) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      mbar_wait(bars + BX_BAR_QFULL, 0);
      tc_fence_after();
      const unsigned long long dq = umma_desc_sw128(base + BX_OFF_Q);   // the blocks' count tiles are contiguous: N rows
      const u32 idesc = (1u << 4) | ((u32)((BX_N * nqb) >> 3) << 17) | ((u32)(BX_M >> 4) << 24);
      const u32 n_acc = (u32)(512 / (BX_N * nqb));
      u32 s = 0, ph = 0, li = 0;
      for (int it = first; it < p.n_items; it += step, ++li) {
        const u32 acc = li % n_acc, aph = (li / n_acc) & 1u;
        mbar_wait(bars + BX_BAR_TEMPTY + 8 * acc, aph ^ 1u);  // the epilogue has drained this accumulator
        mbar_wait(bars + 8 * s, ph);                          // TMA bytes have landed
        tc_fence_after();
        const unsigned long long da = umma_desc_sw128(base + s * BX_A_BYTES);
#pragma unroll
        for (int k = 0; k < BX_K / 16; ++k)  // +32 bytes per K = 16 step inside the swizzle span
          tc_mma_bf16(tmem_base + acc * (u32)(BX_N * nqb), da + 2ull * k, dq + 2ull * k, idesc, k != 0);
        tc_commit(bars + 64 + 8 * s);                 // frees the ring slot when these MMAs retire
        tc_commit(bars + BX_BAR_TFULL + 8 * acc);     // accumulator complete
        if (++s == BX_STAGES) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. +31 =====
    const int e = warp - 2, lq = warp & 3, g = e >> 2;
    // The warp's 32 documents x 32 queries tile of sparse contributions is FIXED POINT (2^-16): shared
    // memory has a native integer atomic add, a float one is a compare-and-swap loop.
    int* S = reinterpret_cast<int*>(gen + BX_OFF_S + e * BX_S_BYTES);
    int4* S4 = reinterpret_cast<int4*>(S) + lane * 8;
    float gm[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) gm[j] = -INFINITY;
    // this warp's items: li = g, g + BX_EPI_GROUPS, ...; an item is served block by block (a "use")
    const long long it_step = (long long)BX_EPI_GROUPS * step;
    const u32 n_acc = (u32)(512 / (BX_N * nqb));
    const size_t bstride = p.bucket_stride;
    // the item's bucket of this lane quarter, per block: blocks are bucket_stride words apart
    auto bucket_of = [&](long long it) -> const u32* {
      return p.buckets + (size_t)qb0 * bstride + ((size_t)it * p.stride * 4 + lq) * BX_CAP;
    };
    // the first 64 words (count + 63 entries) of the buckets of every block of an item, one item ahead
    u32 w0[BX_MAX_QB], w1[BX_MAX_QB];
#pragma unroll
    for (int qb = 0; qb < BX_MAX_QB; ++qb) w0[qb] = w1[qb] = 0;
    long long it = (long long)first + (long long)g * step;
    if (it < p.n_items) {
      const u32* bp = bucket_of(it);
#pragma unroll
      for (int qb = 0; qb < BX_MAX_QB; ++qb)
        if (qb < nqb) {
          w0[qb] = ldg_stream_word(bp + qb * bstride + lane);
          w1[qb] = ldg_stream_word(bp + qb * bstride + 32 + lane);
        }
    }
    for (u32 li = (u32)g; it < p.n_items; it += it_step, li += BX_EPI_GROUPS) {
      const u32* bp = bucket_of(it);
      const long long row = it * p.stride * BX_M + lq * 32 + lane;   // this thread's document
      const u32 acc = li % n_acc, aph = (li / n_acc) & 1u;
      const u32 tbase = tmem_base + ((u32)(lq * 32) << 16) + acc * (u32)(BX_N * nqb);
      // the next item's bucket words are in flight while this item is processed
      u32 n0[BX_MAX_QB], n1[BX_MAX_QB];
      {
        const bool more = it + it_step < p.n_items;
        const u32* nb = bucket_of(it + it_step);
#pragma unroll
        for (int qb = 0; qb < BX_MAX_QB; ++qb) {
          n0[qb] = n1[qb] = 0;
          if (more && qb < nqb) {
            n0[qb] = ldg_stream_word(nb + qb * bstride + lane);
            n1[qb] = ldg_stream_word(nb + qb * bstride + 32 + lane);
          }
        }
      }
      // When the accumulator is already complete (several query blocks per item: the epilogue is the
      // bottleneck) every block's TMEM read is issued before its scatter, which hides the read latency;
      // when it is not (one block: the kernel waits for HBM) the first scatter fills the wait.
      bool ready = __all_sync(0xFFFFFFFFu, mbar_test(bars + BX_BAR_TFULL + 8 * acc, aph));
      if (ready) tc_fence_after();
#pragma unroll
      for (int qb = 0; qb < BX_MAX_QB; ++qb) {
        if (qb >= nqb) break;
        u32 vr[32];
        if (ready) tmem_ld32_issue(tbase + (u32)(qb * BX_N), vr);
        // ---- scatter the bucket's sparse contributions into the warp's tile ----
        int cnt = (int)__shfl_sync(0xFFFFFFFFu, w0[qb], 0);
        cnt = cnt < BX_CAP - 1 ? cnt : BX_CAP - 1;
        {
          auto apply = [&](u32 w) {
            const u32 r = w & 31u, q = (w >> 5) & 31u;
            const float val = __half2float(__ushort_as_half((unsigned short)(w >> 16)));
            atomicAdd(S + r * 32 + ((((q >> 2) ^ (r & 7u))) << 2) + (q & 3u), __float2int_rn(val * 65536.f));
          };
          if (lane >= 1 && lane - 1 < cnt) apply(w0[qb]);
          if (31 + lane < cnt) apply(w1[qb]);
          if (cnt > 63)
            for (int x = 63 + lane; x < cnt; x += 32) apply(bp[qb * bstride + 1 + x]);
        }
        if (!ready) {
          mbar_wait(bars + BX_BAR_TFULL + 8 * acc, aph);
          tc_fence_after();
          ready = true;
          tmem_ld32_issue(tbase + (u32)(qb * BX_N), vr);
        }
        float4 bb[8];   // the block's admission bounds (broadcast reads), in flight with the rest
        if (MODE == BX_MAIN) {
          const float4* b4 = reinterpret_cast<const float4*>(s_thr + qb * BX_N);
#pragma unroll
          for (int c = 0; c < 8; ++c) bb[c] = b4[c];
        }
        __syncwarp();
        tmem_ld32_wait(vr);
        if (qb == nqb - 1) {   // the item's accumulator is in registers (this warp's share): release it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + BX_BAR_TEMPTY + 8 * acc);
        }
        float v[32];
        // ---- add this document's row of the tile (16-byte chunks XOR-swizzled by the row) and clear it ----
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          int4* pp = S4 + (c ^ (lane & 7));
          const int4 s4 = *pp;
          *pp = make_int4(0, 0, 0, 0);
          v[4 * c] = fmaf(__int2float_rn(s4.x), 1.0f / 65536.f, __uint_as_float(vr[4 * c]));
          v[4 * c + 1] = fmaf(__int2float_rn(s4.y), 1.0f / 65536.f, __uint_as_float(vr[4 * c + 1]));
          v[4 * c + 2] = fmaf(__int2float_rn(s4.z), 1.0f / 65536.f, __uint_as_float(vr[4 * c + 2]));
          v[4 * c + 3] = fmaf(__int2float_rn(s4.w), 1.0f / 65536.f, __uint_as_float(vr[4 * c + 3]));
        }
        __syncwarp();
        if (MODE == BX_SAMPLE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) gm[j] = fmaxf(gm[j], v[j]);
        } else {
          bool any0 = false, any1 = false, any2 = false, any3 = false;   // four short chains instead of one long one
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            any0 |= v[4 * c] >= bb[c].x;
            any1 |= v[4 * c + 1] >= bb[c].y;
            any2 |= v[4 * c + 2] >= bb[c].z;
            any3 |= v[4 * c + 3] >= bb[c].w;
          }
          if (((any0 | any1) | (any2 | any3)) && row < p.n_docs) {
            // Rare (about 16 * KP documents per query over the whole pass).
            u32 hits = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) hits |= (v[j] >= s_thr[qb * BX_N + j] ? 1u : 0u) << j;
            const int nq_all = nqb * BX_N;
            while (hits) {
              const int j = __ffs(hits) - 1;
              hits &= hits - 1;
              const int qg = qb * BX_N + j;
              const int slot = atomicAdd(&s_cnt[qg], 1);
              if (slot < p.cap)
                p.cand[((size_t)blockIdx.x * nq_all + qg) * p.cap + slot] = make_key(pick32(v, j), (u32)row);
            }
          }
        }
      }
#pragma unroll
      for (int qb = 0; qb < BX_MAX_QB; ++qb) {
        w0[qb] = n0[qb];
        w1[qb] = n1[qb];
      }
    }
    if (MODE == BX_SAMPLE) {
      // group maxima: the documents an epilogue warp saw form one group (a warp that saw none
      // reports -inf, a valid maximum of nothing); lane j publishes query j's
      float mine = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float m = gm[j];
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, 16));
        if (lane == j) mine = m;
      }
      // the warp's tile is free now: park the 32 maxima there for the cross-group reduction below
      reinterpret_cast<float*>(S)[lane] = mine;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (MODE == BX_SAMPLE && warp >= 2 && warp < 6) {
    // one group of documents per TMEM lane quarter: everything the CTA's epilogue warps of that quarter saw
    // (a quarter that saw nothing reports -inf, a valid maximum of nothing); lane j publishes query j's
    const int e = warp - 2;
    float m = -INFINITY;
#pragma unroll
    for (int g = 0; g < BX_EPI_GROUPS; ++g)
      m = fmaxf(m, reinterpret_cast<const float*>(gen + BX_OFF_S + (e + 4 * g) * BX_S_BYTES)[lane]);
    const int n_groups = (int)gridDim.x * BX_GROUPS_PER_CTA;
    p.gmax[(size_t)(qb0 * BX_N + lane) * n_groups + (size_t)blockIdx.x * BX_GROUPS_PER_CTA + e] = m;
  }
  if (MODE == BX_MAIN)
    for (int i = threadIdx.x; i < nqb * BX_N; i += BX_THREADS) p.cnt[(size_t)blockIdx.x * nqb * BX_N + i] = s_cnt[i];
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Admission bounds: thr[q] = the kp-th largest of the n_groups group maxima gmax[q][*] (attained by
// kp different documents, hence a lower bound of the kp-th best approximate score); -inf with
// fewer than kp groups.  One warp per query: coalesced loads into shared memory, then bit-wise
// bisection on the orderable integer image of the floats (no block barrier).
// ---------------------------------------------------------------------------------------
constexpr int BXT_WARPS = 4;
constexpr int BXT_MAX_GROUPS = 2048;

__global__ void __launch_bounds__(BXT_WARPS * 32)
bm25x_bound_kernel(const float* __restrict__ gmax, int n_groups, int kp, int n_queries, float* __restrict__ thr) {
  __shared__ u32 s_vals[BXT_WARPS][BXT_MAX_GROUPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * BXT_WARPS + warp;
  if (q >= n_queries) return;
  if (n_groups < kp) {
    if (lane == 0) thr[q] = -INFINITY;
    return;
  }
  u32* v = s_vals[warp];
  const float* src = gmax + (size_t)q * n_groups;
  for (int g0 = lane; g0 < n_groups; g0 += 32 * 8) {   // 8 independent loads in flight per lane
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + 32 * u;
      x[u] = g < n_groups ? src[g] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + 32 * u;
      if (g < n_groups) v[g] = f32_orderable(x[u]);
    }
  }
  __syncwarp();
  u32 prefix = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const u32 c = prefix | (1u << bit);
    int local = 0;
    for (int g = lane; g < n_groups; g += 32) local += v[g] >= c;
    if (__reduce_add_sync(0xFFFFFFFFu, local) >= kp) prefix = c;
  }
  if (lane == 0) thr[q] = orderable_f32(prefix);
}

// ---------------------------------------------------------------------------------------
// One CTA per query of the block: KP best candidates -> exact float64 scores -> rank -> certificate.
// ---------------------------------------------------------------------------------------
template <int KPL>
__global__ void __launch_bounds__(FIN_THREADS)
bm25x_finalize_kernel(cmr_lex_index ix, const int* __restrict__ q_terms, const int* __restrict__ q_ptr, int q_base,
                      const u64* __restrict__ cand, const int* __restrict__ cnt, int n_lists, int q_stride, int cap, int cap_total,
                      const float* __restrict__ thr, long long row_offset, int k, double* __restrict__ out_scores,
                      long long* __restrict__ out_ids, int* __restrict__ out_counts, int* __restrict__ out_flags) {
  constexpr int KP = 32 * KPL;
  extern __shared__ __align__(16) unsigned char smem_fin[];
  u64* s_keys = reinterpret_cast<u64*>(smem_fin);          // [cap_total]
  u64* s_out = s_keys + cap_total;                         // [KP]
  double* s_score = reinterpret_cast<double*>(s_out + KP); // [KP]
  u64* s_surv = reinterpret_cast<u64*>(s_score + KP);      // [4 * KP]
  double* s_contrib = reinterpret_cast<double*>(s_surv + 4 * KP);  // [KP][BX_MAXT]
  __shared__ int s_ctl[2];
  __shared__ int s_tok[BX_MAXT];
  __shared__ int s_ntok;
  const int ql = blockIdx.x, q = q_base + ql, tid = threadIdx.x;
  if (out_flags[q] != 0) {  // flagged by the prep / bucket kernels: the exact kernels serve it (CTA-uniform)
    if (tid == 0) out_counts[q] = 0;
    return;
  }
  int n_total, over;
  cand_collect_select<KP>(cand, cnt, n_lists, q_stride, cap, cap_total, ql, s_keys, s_out, s_surv, s_ctl, &n_total, &over);
  if (tid == 0) {
    int n = 0;
    for (int i = q_ptr[q]; i < q_ptr[q + 1]; ++i) {
      const int t = q_terms[i];
      if (t >= 0 && t < ix.n_terms && n < BX_MAXT) s_tok[n++] = t;
    }
    s_ntok = n;
  }
  __syncthreads();
  const int n_valid = count_valid(s_out, KP);
  const int ntok = s_ntok;
  // exact contribution of every (candidate, token): bisection in the term's slice of the document's tile
  for (int idx = tid; idx < n_valid * BX_MAXT; idx += FIN_THREADS) {
    const int c = idx / BX_MAXT, j = idx - c * BX_MAXT;
    double contrib = 0.0;
    if (j < ntok) {
      const int t = s_tok[j];
      const u32 row = key_row(s_out[c]);
      const u32 tile = row / (u32)ix.tile_docs, local = row - tile * (u32)ix.tile_docs;
      const long long tbase = ix.term_ptr[t];
      u32 lo = 0, end = 0;
      lex_slice(ix, t, lex_skip_row(ix, t), (int)tile, &lo, &end);
      u32 hi = end;
      while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if ((ix.post_pack[tbase + mid] & 0xFFFFu) < local) lo = mid + 1;
        else hi = mid;
      }
      if (lo < end) {
        const u32 pk = ix.post_pack[tbase + lo];
        if ((pk & 0xFFFFu) == local) contrib = __dmul_rn(ix.idf[t], ix.imp_table[pk >> 16]);  // rounded product, as rank_bm25
      }
    }
    s_contrib[idx] = contrib;
  }
  __syncthreads();
  if (tid < n_valid) {
    double s = 0.0;
    for (int j = 0; j < ntok; ++j) s = __dadd_rn(s, s_contrib[tid * BX_MAXT + j]);  // query-token order; x + 0.0 == x
    s_score[tid] = s;
  }
  __syncthreads();
  const int n_out = n_valid < k ? n_valid : k;
  // every document outside s_out has approximate score <= cutoff
  double cutoff;
  if (n_total > KP) cutoff = (double)key_score(s_out[KP - 1]);
  else cutoff = (double)thr[ql];   // every candidate was selected; the rest stayed below the admission bound
  const bool open_ended = !(cutoff > -INFINITY);   // no bound at all: every document was a candidate
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
      out_scores[(size_t)q * k + rank] = s;
      out_ids[(size_t)q * k + rank] = (long long)r + row_offset;
    }
    if (rank == n_out - 1) {
      int flag = over ? (CMR_FLAG_UNCERTIFIED | BX_F_LIST) : 0;
      if (!open_ended) {
        const double reach = (cutoff + BX_ABS) / (1.0 - BX_RHO);   // the best exact score an outsider can have
        if (n_out < k || !(s > reach)) flag |= CMR_FLAG_UNCERTIFIED | BX_F_CERT;
      }
      out_flags[q] = flag;
    }
  }
  for (int i = n_out + tid; i < k; i += FIN_THREADS) {
    out_scores[(size_t)q * k + i] = 0.0;
    out_ids[(size_t)q * k + i] = -1;
  }
  if (tid == 0) {
    out_counts[q] = n_out;
    if (n_out == 0) out_flags[q] = (over || !open_ended) ? (CMR_FLAG_UNCERTIFIED | BX_F_CERT) : 0;
  }
}

// ---- host side ------------------------------------------------------------------------
struct BxPlan {
  int kpl, kp, cap, cap_total;
  int n_blocks, n_items, n_sample, stride, grid_main, grid_sample, n_groups, n_b, group_qb;
  size_t bucket_stride;
  size_t off_qmat, off_pairs, off_npairs, off_thr, off_gmax, off_cand, off_cnt, off_buckets, total;
  size_t smem_fin;
};

static void bx_plan(const cmr_lex_index& ix, int n_queries, int k, int sms, BxPlan* p) {
  const int need = k + CMR_SLACK;
  p->kpl = need <= 32 ? 1 : (need <= 64 ? 2 : 4);
  p->kp = 32 * p->kpl;
  p->cap = BX_LIST_CAP;
  p->cap_total = BX_CAP_PER_KP * p->kp;
  p->n_blocks = (n_queries + BX_N - 1) / BX_N;
  p->n_items = (int)((ix.n_docs + BX_M - 1) / BX_M);
  const int full = (int)(ix.n_docs / BX_M);
  int stride = full / BX_SAMPLE_STRIDE;   // small indexes: sample (nearly) every tile
  if (stride < 1) stride = 1;
  if (stride > BX_SAMPLE_STRIDE) stride = BX_SAMPLE_STRIDE;
  p->stride = stride;
  p->n_sample = full > 0 ? (full + stride - 1) / stride : 0;
  p->grid_main = p->n_items < sms ? p->n_items : sms;
  p->grid_sample = p->n_sample < sms ? p->n_sample : sms;
  if (p->grid_sample > BXT_MAX_GROUPS / BX_GROUPS_PER_CTA) p->grid_sample = BXT_MAX_GROUPS / BX_GROUPS_PER_CTA;
  if (p->grid_sample < 1) p->grid_sample = 1;
  p->n_groups = p->grid_sample * BX_GROUPS_PER_CTA;
  if (p->n_groups < p->kp) {
    // no bound (tiny index): every document of a CTA's tiles is a candidate of every query
    const int per_cta = BX_M * ((p->n_items + p->grid_main - 1) / p->grid_main);
    if (per_cta > p->cap) p->cap = per_cta;
  }
  p->n_b = ix.tile_docs / 32;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t off = 0;
  p->off_qmat = off;    off += up((size_t)p->n_blocks * BX_N * BX_K * 2);
  p->off_pairs = off;   off += up((size_t)p->n_blocks * BX_MAX_PAIRS * sizeof(BxPair));
  p->off_npairs = off;  off += up((size_t)p->n_blocks * 4);
  p->group_qb = p->n_blocks < BX_MAX_QB ? p->n_blocks : BX_MAX_QB;   // query blocks served by one pass
  p->bucket_stride = (size_t)ix.n_tiles * p->n_b * BX_CAP;
  const size_t gq = (size_t)p->group_qb * BX_N;
  p->off_thr = off;     off += up(gq * 4);
  p->off_gmax = off;    off += up((size_t)p->n_groups * gq * 4);
  p->off_cand = off;    off += up((size_t)p->grid_main * gq * p->cap * 8);
  p->off_cnt = off;     off += up((size_t)p->grid_main * gq * 4);
  p->off_buckets = off; off += up((size_t)p->group_qb * p->bucket_stride * 4);
  p->total = off;
  p->smem_fin = (size_t)p->cap_total * 8 + (size_t)p->kp * 16 + (size_t)4 * p->kp * 8 + (size_t)p->kp * BX_MAXT * 8 + 16;
}

bool bm25_head_eligible(const cmr_lex_index& ix, int k, bool has_mask) {
  return !has_mask && ix.head_mat != nullptr && ix.head_slot != nullptr && ix.post_pack != nullptr &&
         ix.imp_table != nullptr && ix.n_terms > 0 && ix.n_head >= 0 && ix.n_head <= BX_K && ix.n_docs >= BX_M &&
         ix.n_docs < 0x7FFFFF00ll && ix.tile_docs <= BX_MAX_TILE_DOCS && ix.tile_docs % BX_M == 0 &&
         ((uintptr_t)ix.head_mat % 16) == 0 && k >= 1 && k + CMR_SLACK <= 128;
}

size_t bm25_head_workspace_bytes(const cmr_lex_index& ix, int n_queries, int k) {
  BxPlan p;
  const int sms = sm_count();
  bx_plan(ix, n_queries, k, sms > 0 ? sms : 148, &p);
  return p.total;
}

static int bx_opt_in() {
  static int attr_dev_mask = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_dev_mask & (1 << dev))) {
    cudaError_t e = cudaFuncSetAttribute(bm25x_mma_kernel<BX_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BX_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(bm25x_mma_kernel<BX_MAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BX_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(bm25x_finalize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(bm25x_finalize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(bm25x_finalize_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(bm25x)");
    attr_dev_mask |= (1 << dev);
  }
  return CMR_OK;
}

int bm25_head_topk(const cmr_lex_index& ix, const int* q_terms, const int* q_ptr, int n_queries, int k,
                   long long row_offset, double* out_scores, long long* out_ids, int* out_counts, int* out_flags,
                   void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int sms = sm_count();
  if (sms <= 0) return CMR_ECUDA;
  BxPlan p;
  bx_plan(ix, n_queries, k, sms, &p);
  if (!workspace || workspace_bytes < p.total) {
    set_error("bm25 head path: workspace too small: %zu < %zu", workspace_bytes, p.total);
    return CMR_EWORKSPACE;
  }
  if (p.smem_fin > 200 * 1024) {
    set_error("bm25 head path: finalize shared memory %zu too large", p.smem_fin);
    return CMR_EUNSUPPORTED;
  }
  int rc = bx_opt_in();
  if (rc != CMR_OK) return rc;
  unsigned char* ws = (unsigned char*)workspace;
  __half* qmat = (__half*)(ws + p.off_qmat);
  BxPair* pairs = (BxPair*)(ws + p.off_pairs);
  int* n_pairs = (int*)(ws + p.off_npairs);
  float* thr = (float*)(ws + p.off_thr);
  float* gmax = (float*)(ws + p.off_gmax);
  u64* cand = (u64*)(ws + p.off_cand);
  int* cnt = (int*)(ws + p.off_cnt);
  u32* buckets = (u32*)(ws + p.off_buckets);

  alignas(64) CUtensorMap tm_rows, tm_q;
  rc = make_tmap(&tm_rows, ix.head_mat, ix.n_docs, BX_K, BX_M, true);
  if (rc != CMR_OK) return rc;
  bm25x_prep_kernel<<<p.n_blocks, BX_N * 16, 0, st>>>(ix, q_terms, q_ptr, n_queries, qmat, pairs, n_pairs, out_flags);

  // groups of up to BX_MAX_QB blocks of 32 queries: head_mat streams once per group
  for (int blk0 = 0; blk0 < p.n_blocks; blk0 += p.group_qb) {
    const int n_qb = p.n_blocks - blk0 < p.group_qb ? p.n_blocks - blk0 : p.group_qb;
    const int q_base = blk0 * BX_N;
    const int nq = n_queries - q_base < n_qb * BX_N ? n_queries - q_base : n_qb * BX_N;
    rc = make_tmap(&tm_q, qmat + (size_t)blk0 * BX_N * BX_K, (long long)n_qb * BX_N, BX_K, BX_N, true);
    if (rc != CMR_OK) return rc;
    bm25x_bucket_kernel<<<dim3(ix.n_tiles, n_qb), BXB_THREADS, 0, st>>>(ix, pairs + (size_t)blk0 * BX_MAX_PAIRS, n_pairs + blk0,
                                                                        q_base, buckets, p.bucket_stride, out_flags);
    BxParams kp{};
    kp.n_docs = ix.n_docs;
    kp.n_qb = n_qb;
    kp.n_queries = nq;
    kp.buckets = buckets;
    kp.bucket_stride = p.bucket_stride;
    kp.gmax = gmax;
    kp.thr = thr;
    kp.cand = cand;
    kp.cnt = cnt;
    kp.cap = p.cap;
    if (p.n_sample > 0) {
      kp.n_items = p.n_sample;
      kp.stride = p.stride;
      bm25x_mma_kernel<BX_SAMPLE><<<dim3(p.grid_sample, n_qb), BX_THREADS, BX_SMEM, st>>>(tm_q, tm_rows, kp);
    }
    bm25x_bound_kernel<<<(nq + BXT_WARPS - 1) / BXT_WARPS, BXT_WARPS * 32, 0, st>>>(gmax, p.n_sample > 0 ? p.n_groups : 0,
                                                                                   p.kp, nq, thr);
    kp.n_items = p.n_items;
    kp.stride = 1;
    bm25x_mma_kernel<BX_MAIN><<<p.grid_main, BX_THREADS, BX_SMEM, st>>>(tm_q, tm_rows, kp);
    const int q_stride = n_qb * BX_N;
    switch (p.kpl) {
      case 1:
        bm25x_finalize_kernel<1><<<nq, FIN_THREADS, p.smem_fin, st>>>(ix, q_terms, q_ptr, q_base, cand, cnt, p.grid_main,
                                                                     q_stride, p.cap, p.cap_total, thr, row_offset, k,
                                                                     out_scores, out_ids, out_counts, out_flags);
        break;
      case 2:
        bm25x_finalize_kernel<2><<<nq, FIN_THREADS, p.smem_fin, st>>>(ix, q_terms, q_ptr, q_base, cand, cnt, p.grid_main,
                                                                     q_stride, p.cap, p.cap_total, thr, row_offset, k,
                                                                     out_scores, out_ids, out_counts, out_flags);
        break;
      default:
        bm25x_finalize_kernel<4><<<nq, FIN_THREADS, p.smem_fin, st>>>(ix, q_terms, q_ptr, q_base, cand, cnt, p.grid_main,
                                                                     q_stride, p.cap, p.cap_total, thr, row_offset, k,
                                                                     out_scores, out_ids, out_counts, out_flags);
        break;
    }
  }
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

}  // namespace cmr
