// dense_mma.cu -- A1 batched: exact dense top-k as a tcgen05/TMA GEMM with a top-k epilogue.
//
// Replaces the hnswlib search behind ChromaVectorStore.query (reference
// rag/retrieval/vector_chroma.py:204-253) for batches of queries.  scores = Q . C^T is
// computed on the 5th-generation tensor cores and never written to memory:
//
//   warp 0   TMA producer: per 64-column K chunk one box of queries ([<=128, 64] bf16) and one
//            box of rows ([256, 64] bf16) land in a 4-stage ring of 128-byte-swizzled shared
//            memory tiles (cp.async.bulk.tensor + mbarrier complete_tx).
//   warp 1   one thread issues tcgen05.mma.cta_group::1.kind::f16, M = 128 (queries),
//            N = 256 (rows), K = 16, both operands K-major from shared memory, fp32
//            accumulators in TMEM (two buffers of 256 columns, so the tensor pipe works on
//            tile i+1 while tile i is read back); tcgen05.commit frees ring slots and
//            publishes accumulators.
//   warps 2-5 epilogue: tcgen05.ld.32x32b gives every thread ONE query (TMEM lane) and 32
//            rows per load.  A thread reduces the 32 scores with max and compares once with
//            its query's admission bound; only the rare survivors are appended to the
//            query's candidate buffer (one global atomic each).
//
// The admission bound makes the single pass exact without keeping sorted lists on chip:
// the same kernel first runs in SAMPLE mode over every 16th (32nd from 2M rows on) row tile and writes the maximum
// score of each sampled group of rows; the KP-th largest of those maxima is attained by at
// least KP different rows, hence it is a lower bound of the KP-th best score of the query,
// and every row of the true top-KP passes `score >= bound` in MAIN mode.  About stride*KP rows
// per query pass; dense_finalize_cand_kernel picks the KP best fp32 keys and hands them to
// the shared exact float64 rescoring tail (dense_common.cuh), so ids, order and scores are
// bit-identical to the scan path and to the oracle.
//
// Work distribution: persistent CTAs (one per SM), CTA c owns row tiles c, c+grid, ...; for
// more than 128 queries the query blocks are the inner loop, so a row tile is fetched from
// HBM once and re-read from L2.  Algorithmic bytes per launch: n_rows * dim * 2 (+1/16 for
// the sample pass); flops 2 * n_queries * n_rows * dim.
#include <cuda.h>
#include <stdlib.h>

#include "mma_common.cuh"

namespace cmr {

constexpr int MM_Q = 128;     // queries per MMA tile (UMMA M, TMEM lanes)
constexpr int MM_R = 256;     // rows per MMA tile (UMMA N, TMEM columns)
constexpr int MM_K = 64;      // bf16 per K chunk: one 128-byte swizzle span
constexpr int MM_STAGES = 4;   // ring slots (at most; MmParams::n_stages of them are used)
constexpr int MM_A_BYTES = MM_Q * MM_K * 2;  // 16 KiB
constexpr int MM_B_BYTES = MM_R * MM_K * 2;  // 32 KiB
constexpr int MM_THREADS = 192;
constexpr int MM_SAMPLE = 0, MM_MAIN = 1, MM_NEARDUP = 2;
constexpr int MM_CAP_PER_KP = 128;   // candidates finalize can collect per query = 128 * KP
constexpr int MM_SAMPLE_STRIDE = 16; // SAMPLE mode visits every 16th full tile ...
constexpr int MM_SAMPLE_STRIDE_LARGE = 32;   // ... every 32nd from MM_LARGE_TILES tiles on (2M rows): the
#ifndef CMR_MM_LARGE_TILES
#define CMR_MM_LARGE_TILES 8192
#endif
constexpr int MM_LARGE_TILES = CMR_MM_LARGE_TILES;   // bound is then still the KP-th of >= 256 tile maxima

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3, M >> 4
constexpr u32 MM_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(MM_R >> 3) << 17) | ((u32)(MM_Q >> 4) << 24);

// ---------------------------------------------------------------------------------------
// Work enumeration shared by the three warp roles.  An item is one (A block of 128 rows of
// the left operand, B tile of 256 rows of the right operand) pair = one TMEM accumulator.
//   MM_SAMPLE / MM_MAIN  left = queries, right = matrix.  outer = row tile (CTA c owns tiles
//       c, c+grid, ...; SAMPLE: every `stride`-th full tile), inner = query block, so a row
//       tile is fetched from HBM once and re-read from L2 for further query blocks.
//   MM_NEARDUP           left = right = matrix, lower triangle only: outer = block of 128
//       rows i (blocks begin, begin+step, ... of this rank, dealt round-robin to the CTAs),
//       inner = every 256-row tile that holds some j < i.
// ---------------------------------------------------------------------------------------
struct MmParams {
  int n_left;          // rows of the left operand (queries; NEARDUP: matrix rows)
  long long n_rows;    // rows of the right operand
  int n_chunks;        // K chunks of 64 columns
  int n_stages;        // ring slots in use (2..MM_STAGES): fewer leave shared memory to a co-resident kernel
  int n_outer;         // outer items in total
  int n_inner;         // SAMPLE/MAIN: query blocks
  int stride;          // SAMPLE: tile stride; NEARDUP: block step of this rank
  int begin;           // NEARDUP: first block of this rank
  u32 tx_bytes;        // bytes per ring stage (both TMA boxes)
  u32 a_bytes;         // shared memory reserved for the left box of a stage (multiple of 1024;
                       // < 16 KiB when fewer than 128 queries: the MMA then reads past it into
                       // the stage's own right tile for lanes nobody looks at)
  int rows_evict_first;
  // SAMPLE / MAIN
  const float* thr;
  float* gmax;
  int gstride, gpt;
  u64* cand;           // [n_ctas][n_left][cap]: every CTA keeps private lists, so no atomics
  int* cnt;            // [n_ctas][n_left] entries appended (may exceed cap: the excess was dropped)
  int cap;             // slots per (CTA, query)
  const u32* mask_bits;  // SAMPLE / MAIN: bit r % 32 of word r / 32 = row r may be returned; NULL = all
  // NEARDUP
  float nd_bound;      // emit pairs with fp32 score >= nd_bound (= threshold - error bound)
  u64* edges;          // (i << 32) | j
  unsigned long long* edge_count;
  unsigned long long edge_cap;
};

template <int MODE>
struct MmIter {
  const MmParams& p;
  int outer_idx, outer, inner, inner_end;
  __device__ __forceinline__ MmIter(const MmParams& pp) : p(pp), outer_idx((int)blockIdx.x), inner(0), inner_end(0) {
    load_outer();
  }
  __device__ __forceinline__ void load_outer() {
    if (outer_idx >= p.n_outer) return;
    if (MODE == MM_NEARDUP) {
      outer = p.begin + outer_idx * p.stride;                  // block of rows i
      inner_end = (outer * MM_Q + MM_Q - 2) / MM_R + 1;        // tiles holding some j <= i_max - 1
    } else {
      outer = outer_idx * (MODE == MM_SAMPLE ? p.stride : 1);  // row tile
      inner_end = p.n_inner;
    }
    inner = 0;
  }
  __device__ __forceinline__ bool valid() const { return outer_idx < p.n_outer; }
  __device__ __forceinline__ void advance() {
    if (++inner >= inner_end) {
      outer_idx += (int)gridDim.x;
      load_outer();
    }
  }
  // first row of the left (A) and right (B) boxes of the current item
  __device__ __forceinline__ int a_row0() const { return (MODE == MM_NEARDUP ? outer : inner) * MM_Q; }
  __device__ __forceinline__ int b_row0() const { return (MODE == MM_NEARDUP ? inner : outer) * MM_R; }
};

// MODE == MM_SAMPLE: the epilogue writes group maxima gmax[(item * gpt + g) * gstride + query],
//   gpt = 8 (groups of 32 rows) or 1 (the whole tile); sampled tiles are always full tiles.
// MODE == MM_MAIN:   scores >= thr[query] are appended to this CTA's private list of the query
//   as (orderable fp32 score, ~row) keys.  A query is always handled by the same thread of a
//   CTA, so the list counters live in shared memory and need no atomics.
// MODE == MM_NEARDUP: pairs (i, j < i) with score >= nd_bound are appended to `edges`.
template <int MODE>
__global__ void __launch_bounds__(MM_THREADS, 1)
dense_mma_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_rows,
                 const MmParams p) {
  extern __shared__ unsigned char smem_raw[];
  const u32 raw = smem_u32(smem_raw);
  const u32 base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const u32 stage_bytes = p.a_bytes + MM_B_BYTES;
  const u32 bars = base + (u32)p.n_stages * stage_bytes;
  // barrier slots: full[s] at +8s, empty[s] at +32+8s, tmem_full[b] at +64+8b, tmem_empty[b] at +80+8b
  volatile u32* tmem_slot = reinterpret_cast<volatile u32*>(smem_raw + (bars - raw) + 96);
  int* s_cnt = reinterpret_cast<int*>(smem_raw + (bars - raw) + 128);  // MAIN: [n_inner * 128] list lengths
  if (MODE == MM_MAIN)
    for (int i = threadIdx.x; i < p.n_inner * MM_Q; i += MM_THREADS) s_cnt[i] = 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_rows) : "memory");
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 32 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + 64 + 8 * b, 1);
      mbar_init(bars + 80 + 8 * b, 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the allocating warp also frees
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + 96), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const unsigned long long hint_rows = p.rows_evict_first ? TMA_EVICT_FIRST : TMA_EVICT_NORMAL;
      const unsigned long long hint_left = MODE == MM_NEARDUP ? TMA_EVICT_NORMAL : TMA_EVICT_LAST;
      u32 s = 0, ph = 0;  // ring slot and its phase
      for (MmIter<MODE> w(p); w.valid(); w.advance()) {
        const int a0 = w.a_row0(), b0 = w.b_row0();
        for (int kc = 0; kc < p.n_chunks; ++kc) {
          mbar_wait(bars + 32 + 8 * s, ph ^ 1u);
          const u32 full = bars + 8 * s;
          mbar_expect_tx(full, p.tx_bytes);
          const u32 sa = base + s * stage_bytes;
          tma_load_2d(sa, &tm_q, full, kc * MM_K, a0, hint_left);
          tma_load_2d(sa + p.a_bytes, &tm_rows, full, kc * MM_K, b0, hint_rows);
          if (++s == (u32)p.n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      u32 s = 0, ph = 0, ai = 0;
      for (MmIter<MODE> w(p); w.valid(); w.advance(), ++ai) {
        const u32 ab = ai & 1u, aph = (ai >> 1) & 1u;
        mbar_wait(bars + 80 + 8 * ab, aph ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const u32 d_tmem = tmem_base + ab * MM_R;
        for (int kc = 0; kc < p.n_chunks; ++kc) {
          mbar_wait(bars + 8 * s, ph);  // TMA bytes have landed
          tc_fence_after();
          const u32 sa = base + s * stage_bytes;
          const unsigned long long da = umma_desc_sw128(sa);
          const unsigned long long db = umma_desc_sw128(sa + p.a_bytes);
#pragma unroll
          for (int k = 0; k < MM_K / 16; ++k)  // +32 bytes per K = 16 step inside the swizzle span
            tc_mma_bf16(d_tmem, da + 2ull * k, db + 2ull * k, MM_IDESC, (kc | k) != 0);
          tc_commit(bars + 32 + 8 * s);  // frees the ring slot when these MMAs retire
          if (++s == (u32)p.n_stages) { s = 0; ph ^= 1u; }
        }
        tc_commit(bars + 64 + 8 * ab);   // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 =====
    const int lg = warp & 3;
    const int lane_row = lg * 32 + lane;
    u32 ai = 0;
    for (MmIter<MODE> w(p); w.valid(); w.advance(), ++ai) {
      const u32 ab = ai & 1u, aph = (ai >> 1) & 1u;
      const long long row0 = w.b_row0();
      const int qi = w.a_row0() + lane_row;     // query (NEARDUP: row i) of this thread
      const bool q_ok = qi < p.n_left;
      float bound = INFINITY;
      if (MODE == MM_MAIN && q_ok) bound = p.thr[qi];
      if (MODE == MM_NEARDUP && q_ok) bound = p.nd_bound;
      mbar_wait(bars + 64 + 8 * ab, aph);
      tc_fence_after();
      const u32 taddr = tmem_base + ((u32)(lg * 32) << 16) + ab * MM_R;
      float tile_max = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < MM_R / 32; ++c) {
        __syncwarp();
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        if (MODE != MM_NEARDUP && p.mask_bits != nullptr) {
          // `where` filter / tombstones: one broadcast word covers the 32 rows of this load
          const u32 bits = p.mask_bits[(row0 + c * 32) >> 5];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : -INFINITY;
        }
        float m = v[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
        if (MODE == MM_SAMPLE) {
          if (p.gpt == 8) {
            if (q_ok) p.gmax[((size_t)w.outer_idx * 8 + c) * p.gstride + qi] = m;
          } else {
            tile_max = fmaxf(tile_max, m);
          }
        } else if (q_ok && m >= bound) {
          // Rare per thread (about stride*KP rows per query over the whole scan; near-duplicate
          // pairs in NEARDUP) but at 1024 queries ~40 % of the warp-chunks have a lane in here,
          // and an epilogue warp has no other warp to hide behind: keep it short.  One branch-free
          // pass builds the hit mask, then only the set bits are visited (jump-table pick).
          u32 hits = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) hits |= (v[j] >= bound ? 1u : 0u) << j;
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            const float sv = pick32(v, j);
            const long long row = row0 + c * 32 + j;
            if (MODE == MM_MAIN) {
              if (row < p.n_rows) {
                const int slot = s_cnt[qi]++;   // only this thread touches query qi in this CTA
                if (slot < p.cap)
                  p.cand[((size_t)blockIdx.x * p.n_left + qi) * p.cap + slot] = make_key(sv, (u32)row);
              }
            } else if (row < (long long)qi) {
              const unsigned long long slot = atomicAdd(p.edge_count, 1ull);
              if (slot < p.edge_cap) p.edges[slot] = ((u64)(u32)qi << 32) | (u64)(u32)row;
            }
          }
        }
      }
      if (MODE == MM_SAMPLE && p.gpt != 8 && q_ok) p.gmax[(size_t)w.outer_idx * p.gstride + qi] = tile_max;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 80 + 8 * ab);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (MODE == MM_MAIN)
    for (int i = threadIdx.x; i < p.n_left; i += MM_THREADS) p.cnt[(size_t)blockIdx.x * p.n_left + i] = s_cnt[i];
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// uint8 row mask -> one bit per row, zero padded to whole 256-row tiles (so rows past the end
// of the matrix are filtered too)
__global__ void mask_to_bits_kernel(const uint8_t* __restrict__ mask, long long n_rows, long long n_words,
                                    u32* __restrict__ bits) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  u32 out = 0;
  const long long r0 = w * 32;
  if (r0 + 32 <= n_rows && (reinterpret_cast<uintptr_t>(mask + r0) & 15) == 0) {
    const uint4 a = *reinterpret_cast<const uint4*>(mask + r0), b = *reinterpret_cast<const uint4*>(mask + r0 + 16);
    const u32 words[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) out |= (((words[i] >> (8 * j)) & 0xFFu) ? 1u : 0u) << (4 * i + j);
  } else {
    for (int j = 0; j < 32 && r0 + j < n_rows; ++j) out |= (mask[r0 + j] ? 1u : 0u) << j;
  }
  bits[w] = out;
}

// One CTA per query: bound = the kp-th largest of the G sampled group maxima (bit-wise
// bisection on the orderable integer image of the floats), -inf when there are fewer than
// kp groups.
__global__ void __launch_bounds__(256)
dense_thresh_kernel(const float* __restrict__ gmax, int n_groups, int gstride, int kp,
                    float* __restrict__ thr) {
  extern __shared__ u32 s_vals[];
  __shared__ int s_count;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (n_groups < kp) {
    if (tid == 0) thr[q] = -INFINITY;
    return;
  }
  for (int g = tid; g < n_groups; g += 256) s_vals[g] = f32_orderable(gmax[(size_t)g * gstride + q]);
  if (tid == 0) s_count = 0;
  __syncthreads();
  u32 prefix = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const u32 c = prefix | (1u << bit);
    int local = 0;
    for (int g = tid; g < n_groups; g += 256) local += s_vals[g] >= c;
    local = __reduce_add_sync(0xFFFFFFFFu, local);
    if ((tid & 31) == 0 && local) atomicAdd(&s_count, local);
    __syncthreads();
    if (s_count >= kp) prefix = c;
    __syncthreads();
    if (tid == 0) s_count = 0;
    __syncthreads();
  }
  if (tid == 0) thr[q] = orderable_f32(prefix);
}

// Same bound, one WARP per query (8 queries per CTA) for up to 4096 groups: the maxima are
// staged in the warp's slice of shared memory (conflict-free lane-strided reads) and the 32
// bisection rounds need no block barrier -- this sits on the critical path of every batch.
constexpr int THR_WARPS = 8;
constexpr int THR_WARP_MAX_GROUPS = 4096;

__global__ void __launch_bounds__(THR_WARPS * 32)
dense_thresh_warp_kernel(const float* __restrict__ gmax, int n_groups, int gstride, int kp, int n_queries,
                         float* __restrict__ thr) {
  extern __shared__ u32 s_vals[];  // [THR_WARPS][n_groups]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * THR_WARPS + warp;
  if (q >= n_queries) return;
  if (n_groups < kp) {
    if (lane == 0) thr[q] = -INFINITY;
    return;
  }
  u32* v = s_vals + (size_t)warp * n_groups;
  for (int g0 = lane; g0 < n_groups; g0 += 32 * 8) {  // 8 independent loads in flight per lane
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + 32 * u;
      if (g < n_groups) x[u] = gmax[(size_t)g * gstride + q];
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

// One CTA per query: collect the query's candidates from every CTA's private list, take the KP
// best (unique keys, ranked by counting), then the shared exact rescoring tail.  A list or the
// collection buffer that overflowed -> CMR_FLAG_UNCERTIFIED.
template <int KPL>
__global__ void __launch_bounds__(FIN_THREADS)
dense_finalize_cand_kernel(const u64* __restrict__ cand, const int* __restrict__ cnt, int n_lists, int n_queries,
                           int cap, int cap_total, const uint16_t* __restrict__ emb, int dim,
                           const uint16_t* __restrict__ queries, long long row_offset, int k, double cert_eps,
                           double* __restrict__ out_scores, long long* __restrict__ out_ids,
                           int* __restrict__ out_counts, int* __restrict__ out_flags) {
  constexpr int KP = 32 * KPL;
  extern __shared__ __align__(16) unsigned char smem_fin[];
  u64* s_keys = reinterpret_cast<u64*>(smem_fin);  // [cap_total]
  u64* s_out = s_keys + cap_total;                 // [KP]
  double* s_score = reinterpret_cast<double*>(s_out + KP);
  u64* s_surv = reinterpret_cast<u64*>(s_score + KP);   // [4 * KP]
  __shared__ int s_ctl[2];
  const int qi = blockIdx.x;
  int n_total, s_over;
  cand_collect_select<KP>(cand, cnt, n_lists, n_queries, cap, cap_total, qi, s_keys, s_out, s_surv, s_ctl,
                          &n_total, &s_over);
  dense_finalize_tail<KP>(s_out, s_score, emb, dim, queries + (size_t)qi * dim, row_offset, k, cert_eps,
                          s_over ? CMR_FLAG_UNCERTIFIED : 0, qi, out_scores, out_ids, out_counts, out_flags);
}

// ---- host side ------------------------------------------------------------------------
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tmap_encode_fn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* map, const void* ptr, long long n_rows, int dim, int box_rows, bool f16) {
  tmap_encode_fn enc = tmap_encoder();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return CMR_ECUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)n_rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)dim * 2};
  const cuuint32_t box[2] = {(cuuint32_t)MM_K, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%lld, %d] box %d", (int)r, n_rows, dim, box_rows);
    return CMR_ECUDA;
  }
  return CMR_OK;
}

constexpr int MM_SMEM_MAX = 227 * 1024;
constexpr int MM_MAX_QUERIES = 8192;  // list counters of all query blocks must fit in shared memory

// Ring depth: MM_STAGES unless CMR_MM_STAGES (2..MM_STAGES) says otherwise.  The HBM-bound shapes
// (<= 128 queries) keep the memory system busy with 3 slots (96 KB of rows in flight per SM); the
// shared memory saved lets the BM25 tile kernel run beside the scan (engine.HybridEngine overlap).
static int mma_stages() {
  const char* e = getenv("CMR_MM_STAGES");  // read per call: cheap, and a process may change it
  int n = e ? atoi(e) : MM_STAGES;
  if (n < 2 || n > MM_STAGES) n = MM_STAGES;
  return n;
}

static inline size_t mma_smem_bytes(u32 a_bytes, int n_inner) {
  return (size_t)mma_stages() * (a_bytes + MM_B_BYTES) + 1024 /* alignment slack */ + 128 /* barriers */ +
         (size_t)n_inner * MM_Q * 4 /* MAIN: list counters */;
}

struct MmaPlan {
  int kpl, kp;
  int cap;        // slots of one (CTA, query) list
  int cap_total;  // candidates finalize can collect per query
  int n_lists;    // CTAs of the MAIN pass
  int n_mb, n_chunks, q_box_rows, bpad;
  int n_tiles, n_sample, sample_stride, gpt, n_groups;
  size_t off_cnt, off_thr, off_gmax, off_bits, total;
};

static void mma_plan(long long n_rows, int dim, int n_queries, int k, int sms, MmaPlan* p) {
  const int need = k + CMR_SLACK;
  p->kpl = need <= 32 ? 1 : (need <= 64 ? 2 : 4);
  p->kp = 32 * p->kpl;
  p->cap_total = MM_CAP_PER_KP * p->kp;
  p->n_mb = (n_queries + MM_Q - 1) / MM_Q;
  p->n_chunks = (dim + MM_K - 1) / MM_K;
  p->q_box_rows = n_queries >= MM_Q ? MM_Q : (n_queries + 7) / 8 * 8;
  p->bpad = (n_queries + 31) / 32 * 32;
  p->n_tiles = (int)((n_rows + MM_R - 1) / MM_R);
  const int full_tiles = (int)(n_rows / MM_R);
  const int max_stride = full_tiles >= MM_LARGE_TILES ? MM_SAMPLE_STRIDE_LARGE : MM_SAMPLE_STRIDE;
  int stride = full_tiles / MM_SAMPLE_STRIDE;  // small matrices: sample (nearly) every tile
  if (stride < 1) stride = 1;
  if (stride > max_stride) stride = max_stride;
  for (;;) {
    p->sample_stride = stride;
    p->n_sample = full_tiles > 0 ? (full_tiles + stride - 1) / stride : 0;
    p->gpt = p->n_sample < 8 * p->kp ? 8 : 1;
    p->n_groups = p->n_sample * p->gpt;
    if (p->n_groups <= MM_MAX_GROUPS) break;
    stride *= 2;
  }
  p->n_lists = p->n_tiles < sms ? p->n_tiles : sms;
  if (p->n_groups < p->kp) {
    // no bound (tiny matrix): every row of a CTA's tiles is a candidate
    p->cap = MM_R * ((p->n_tiles + p->n_lists - 1) / p->n_lists);
  } else {
    // About stride * KP candidates per query in all.  An even spread would put only a
    // few into each CTA's list, but neighbouring rows are often similar (chunks of one document)
    // and land in the same tile, so lists are as deep as a 256 MB budget allows, 32..256 slots.
    const long long budget = (256ll << 20) / ((long long)p->n_lists * n_queries * (long long)sizeof(u64));
    p->cap = (int)(budget < 32 ? 32 : (budget > 256 ? 256 : budget));
  }
  if (p->cap > p->cap_total) p->cap = p->cap_total;
  size_t off = ((size_t)p->n_lists * n_queries * p->cap * sizeof(u64) + 15) / 16 * 16;
  p->off_cnt = off;
  off += ((size_t)p->n_lists * n_queries * 4 + 15) / 16 * 16;
  p->off_thr = off;
  off += ((size_t)p->bpad * 4 + 15) / 16 * 16;
  p->off_gmax = off;
  off += ((size_t)(p->n_groups > 0 ? p->n_groups : 1) * p->bpad * 4 + 15) / 16 * 16;
  p->off_bits = off;
  off += (size_t)p->n_tiles * (MM_R / 32) * 4;   // row-mask bits (used when the call has a mask)
  p->total = off;
}

bool dense_mma_eligible(long long n_rows, int dim, int n_queries, int k, bool has_mask) {
  (void)has_mask;  // filters are applied in the epilogue from a bitmask
  return dim >= MM_K && n_rows >= MM_R && n_rows < 0x7FFFFF00ll && n_queries >= 1 &&
         n_queries <= MM_MAX_QUERIES && k >= 1 && k <= CMR_MAX_K;
}

size_t dense_mma_workspace_bytes(long long n_rows, int dim, int n_queries, int k) {
  MmaPlan p;
  const int sms = sm_count();
  mma_plan(n_rows, dim, n_queries, k, sms > 0 ? sms : 148, &p);
  return p.total;
}

template <int KPL>
static int launch_finalize_cand(const DenseArgs& a, const MmaPlan& p, const u64* cand, const int* cnt) {
  const size_t smem = (size_t)p.cap_total * 8 + (size_t)p.kp * 16 + (size_t)4 * p.kp * 8 + 16;
  static int attr_dev_mask = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_dev_mask & (1 << dev))) {
    cudaError_t e = cudaFuncSetAttribute(dense_finalize_cand_kernel<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MM_CAP_PER_KP * 32 * KPL * 8 + 32 * KPL * 16 + 4 * 32 * KPL * 8 + 16);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(dense_finalize_cand)");
    attr_dev_mask |= (1 << dev);
  }
  dense_finalize_cand_kernel<KPL><<<a.n_queries, FIN_THREADS, smem, a.stream>>>(
      cand, cnt, p.n_lists, a.n_queries, p.cap, p.cap_total, a.emb, a.dim, a.queries, a.row_offset, a.k, a.cert_eps,
      a.out_scores, a.out_ids, a.out_counts, a.out_flags);
  return CMR_OK;
}

// one-time (per device) opt-in of the GEMM kernels to their dynamic shared memory
static int mma_opt_in() {
  static int attr_dev_mask = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_dev_mask & (1 << dev))) {
    cudaError_t e = cudaFuncSetAttribute(dense_mma_kernel<MM_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_mma_kernel<MM_MAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_mma_kernel<MM_NEARDUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_thresh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_MAX_GROUPS * 4);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_thresh_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               THR_WARPS * THR_WARP_MAX_GROUPS * 4);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(dense_mma)");
    attr_dev_mask |= (1 << dev);
  }
  return CMR_OK;
}

int launch_admission_bound(const float* gmax, int n_groups, int gstride, int kp, int n_queries, float* thr,
                           cudaStream_t st) {
  const int rc = mma_opt_in();
  if (rc != CMR_OK) return rc;
  if (n_groups > MM_MAX_GROUPS) {
    set_error("admission bound: %d groups exceed %d", n_groups, MM_MAX_GROUPS);
    return CMR_EUNSUPPORTED;
  }
  if (n_groups <= THR_WARP_MAX_GROUPS)
    dense_thresh_warp_kernel<<<(n_queries + THR_WARPS - 1) / THR_WARPS, THR_WARPS * 32,
                               (size_t)THR_WARPS * (n_groups > 0 ? n_groups : 1) * 4, st>>>(
        gmax, n_groups, gstride, kp, n_queries, thr);
  else
    dense_thresh_kernel<<<n_queries, 256, (size_t)n_groups * 4, st>>>(gmax, n_groups, gstride, kp, thr);
  return CMR_OK;
}

int dense_mma_topk(const DenseArgs& a) {
  MmaPlan p;
  const int sms = sm_count();
  if (sms <= 0) return CMR_ECUDA;
  mma_plan(a.n_rows, a.dim, a.n_queries, a.k, sms, &p);
  if (!a.workspace || a.workspace_bytes < p.total) {
    set_error("workspace too small: %zu < %zu", a.workspace_bytes, p.total);
    return CMR_EWORKSPACE;
  }
  CMR_CHECK_ARG(((uintptr_t)a.workspace % 16) == 0, "workspace must be 16-byte aligned");
  { const int rc_attr = mma_opt_in(); if (rc_attr != CMR_OK) return rc_attr; }
  alignas(64) CUtensorMap tm_q, tm_rows;
  int rc = make_tmap(&tm_q, a.queries, a.n_queries, a.dim, p.q_box_rows);
  if (rc != CMR_OK) return rc;
  rc = make_tmap(&tm_rows, a.emb, a.n_rows, a.dim, MM_R);
  if (rc != CMR_OK) return rc;

  unsigned char* ws = (unsigned char*)a.workspace;
  u64* cand = (u64*)ws;
  int* cnt = (int*)(ws + p.off_cnt);
  float* thr = (float*)(ws + p.off_thr);
  float* gmax = (float*)(ws + p.off_gmax);
  const u32 tx_bytes = (u32)(p.q_box_rows * MM_K * 2 + MM_B_BYTES);

  MmParams kp{};
  kp.n_left = a.n_queries;
  kp.n_rows = a.n_rows;
  kp.n_chunks = p.n_chunks;
  kp.n_stages = mma_stages();
  kp.n_inner = p.n_mb;
  kp.tx_bytes = tx_bytes;
  kp.a_bytes = (u32)(p.q_box_rows * MM_K * 2 + 1023) / 1024 * 1024;
  const size_t smem_bytes = mma_smem_bytes(kp.a_bytes, p.n_mb);
  if (smem_bytes > (size_t)MM_SMEM_MAX) {
    set_error("tcgen05 path: %d queries need %zu bytes of shared memory", a.n_queries, smem_bytes);
    return CMR_EUNSUPPORTED;
  }
  kp.thr = thr;
  kp.gmax = gmax;
  kp.gstride = p.bpad;
  kp.gpt = p.gpt;
  kp.cand = cand;
  kp.cnt = cnt;
  kp.cap = p.cap;
  kp.mask_bits = nullptr;
  if (a.row_mask != nullptr) {
    u32* bits = (u32*)(ws + p.off_bits);
    const long long n_words = (long long)p.n_tiles * (MM_R / 32);
    mask_to_bits_kernel<<<(int)((n_words + 255) / 256), 256, 0, a.stream>>>(a.row_mask, a.n_rows, n_words, bits);
    kp.mask_bits = bits;
  }
  if (p.n_sample > 0) {
    const int grid = p.n_sample < sms ? p.n_sample : sms;
    kp.n_outer = p.n_sample;
    kp.stride = p.sample_stride;
    kp.rows_evict_first = 0;
    dense_mma_kernel<MM_SAMPLE><<<grid, MM_THREADS, smem_bytes, a.stream>>>(tm_q, tm_rows, kp);
  }
  rc = launch_admission_bound(gmax, p.n_groups, p.bpad, p.kp, a.n_queries, thr, a.stream);
  if (rc != CMR_OK) return rc;
  {
    const int grid = p.n_lists;
    kp.n_outer = p.n_tiles;
    kp.stride = 1;
    kp.rows_evict_first = p.n_mb == 1;
    dense_mma_kernel<MM_MAIN><<<grid, MM_THREADS, smem_bytes, a.stream>>>(tm_q, tm_rows, kp);
  }
  switch (p.kpl) {
    case 1: rc = launch_finalize_cand<1>(a, p, cand, cnt); break;
    case 2: rc = launch_finalize_cand<2>(a, p, cand, cnt); break;
    default: rc = launch_finalize_cand<4>(a, p, cand, cnt); break;
  }
  if (rc != CMR_OK) return rc;
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

// ---- A9 / K6: near-duplicate filter ---------------------------------------------------------
// (extension; greedy keep-first rule of rag/utils/dedup.py:40-55 with its '>=' comparison,
//  hook rag/admin/backup.py:226-233)
//
// 1. dense_mma_kernel<MM_NEARDUP>: C . C^T on the tensor cores over the lower triangle, pairs
//    (i, j < i) with fp32 score >= threshold - error bound go to the edge buffer.
// 2. neardup_rescore_kernel: one warp per candidate pair recomputes the exact float64 dot
//    (pinned order) and keeps the pair iff exact >= threshold; survivors are compacted.
// 3. the caller sorts the surviving edges by (i, j) (they are 64-bit keys i << 32 | j).
// 4. neardup_resolve_kernel: keep[i] = no edge (i, j) with keep[j]; rows in ascending order, one
//    warp, lanes share the edges of a row.

__global__ void __launch_bounds__(256)
neardup_rescore_kernel(const uint16_t* __restrict__ emb, int dim, const u64* __restrict__ edges,
                       unsigned long long n_edges, double threshold, u64* __restrict__ out_edges,
                       unsigned long long* __restrict__ out_count) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long e = warp0; e < n_edges; e += n_warps) {
    const u64 key = edges[e];
    const u32 i = (u32)(key >> 32), j = (u32)key;
    const double s = warp_exact_dot(emb + (size_t)i * dim, emb + (size_t)j * dim, dim, lane);
    if (lane == 0 && s >= threshold) out_edges[atomicAdd(out_count, 1ull)] = key;
  }
}

__global__ void __launch_bounds__(32)
neardup_resolve_kernel(const u64* __restrict__ sorted_edges, unsigned long long n_edges, long long n_rows,
                       uint8_t* __restrict__ keep) {
  const int lane = threadIdx.x;
  volatile uint8_t* vkeep = keep;  // rows decided earlier in this loop are read back
  for (long long r = lane; r < n_rows; r += 32) keep[r] = 1;
  __syncwarp();
  unsigned long long e = 0;
  while (e < n_edges) {
    const u32 i = (u32)(sorted_edges[e] >> 32);
    bool dup = false;
    unsigned long long e2 = e;
    // edges of row i are contiguous; every j < i is already final
    for (;;) {
      const unsigned long long idx = e2 + lane;
      bool mine = false, same = false;
      if (idx < n_edges) {
        const u64 key = sorted_edges[idx];
        same = (u32)(key >> 32) == i;
        if (same) mine = vkeep[(u32)key] != 0;
      }
      dup |= __any_sync(0xFFFFFFFFu, mine);
      const unsigned ok = __ballot_sync(0xFFFFFFFFu, same);
      e2 += __popc(ok);
      if (ok != 0xFFFFFFFFu) break;
    }
    if (lane == 0 && dup) vkeep[i] = 0;
    __syncwarp();
    __threadfence_block();
    e = e2;
  }
}

}  // namespace cmr

using namespace cmr;

extern "C" int cmr_neardup_edges(const uint16_t* emb, int64_t n_rows, int dim, float bound, int block_begin,
                                 int block_step, uint64_t* out_edges, uint64_t edge_cap, uint64_t* out_count,
                                 cmr_stream_t stream) {
  CMR_CHECK_ARG(emb && out_edges && out_count, "null pointer argument");
  CMR_CHECK_ARG(n_rows >= 2 && n_rows < 0x7FFFFF00ll, "n_rows %lld out of range", (long long)n_rows);
  CMR_CHECK_ARG(dim >= MM_K && dim % 8 == 0 && dim <= 2048, "dim %d must be a multiple of 8 in [64, 2048]", dim);
  CMR_CHECK_ARG(block_begin >= 0 && block_step >= 1, "bad block range");
  CMR_CHECK_ARG(((uintptr_t)emb % 16) == 0, "emb must be 16-byte aligned");
  const int sms = sm_count();
  if (sms <= 0) return CMR_ECUDA;
  int rc = mma_opt_in();
  if (rc != CMR_OK) return rc;
  alignas(64) CUtensorMap tm_left, tm_right;
  rc = make_tmap(&tm_left, emb, n_rows, dim, MM_Q);
  if (rc != CMR_OK) return rc;
  rc = make_tmap(&tm_right, emb, n_rows, dim, MM_R);
  if (rc != CMR_OK) return rc;
  const int n_blocks = (int)((n_rows + MM_Q - 1) / MM_Q);
  const int mine = block_begin < n_blocks ? (n_blocks - block_begin + block_step - 1) / block_step : 0;
  cudaStream_t st = (cudaStream_t)stream;
  CMR_CUDA(cudaMemsetAsync(out_count, 0, sizeof(uint64_t), st));
  if (mine == 0) return CMR_OK;
  MmParams kp{};
  kp.n_left = (int)n_rows;
  kp.n_rows = n_rows;
  kp.n_chunks = (dim + MM_K - 1) / MM_K;
  kp.n_stages = mma_stages();
  kp.n_outer = mine;
  kp.n_inner = 0;
  kp.stride = block_step;
  kp.begin = block_begin;
  kp.tx_bytes = (u32)(MM_A_BYTES + MM_B_BYTES);
  kp.a_bytes = MM_A_BYTES;
  kp.rows_evict_first = 0;
  kp.nd_bound = bound;
  kp.edges = (u64*)out_edges;
  kp.edge_count = (unsigned long long*)out_count;
  kp.edge_cap = edge_cap;
  const int grid = mine < sms ? mine : sms;
  dense_mma_kernel<MM_NEARDUP><<<grid, MM_THREADS, mma_smem_bytes(MM_A_BYTES, 0), st>>>(tm_left, tm_right, kp);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_neardup_rescore(const uint16_t* emb, int dim, const uint64_t* edges, uint64_t n_edges,
                                   double threshold, uint64_t* out_edges, uint64_t* out_count, cmr_stream_t stream) {
  CMR_CHECK_ARG(emb && out_count && (n_edges == 0 || (edges && out_edges)), "null pointer argument");
  CMR_CHECK_ARG(dim > 0 && dim % 8 == 0, "dim must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  CMR_CUDA(cudaMemsetAsync(out_count, 0, sizeof(uint64_t), st));
  if (n_edges == 0) return CMR_OK;
  unsigned long long blocks = (n_edges + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  neardup_rescore_kernel<<<(int)blocks, 256, 0, st>>>(emb, dim, (const u64*)edges, n_edges, threshold,
                                                      (u64*)out_edges, (unsigned long long*)out_count);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}

extern "C" int cmr_neardup_resolve(const uint64_t* sorted_edges, uint64_t n_edges, int64_t n_rows, uint8_t* keep,
                                   cmr_stream_t stream) {
  CMR_CHECK_ARG(n_rows >= 0 && (n_rows == 0 || keep) && (n_edges == 0 || sorted_edges), "bad arguments");
  if (n_rows == 0) return CMR_OK;
  neardup_resolve_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const u64*)sorted_edges, n_edges, n_rows, keep);
  CMR_CUDA(cudaGetLastError());
  return CMR_OK;
}
