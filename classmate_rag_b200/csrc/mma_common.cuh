// mma_common.cuh -- pieces shared by the tcgen05/TMA kernels (dense_mma.cu: batched dense top-k and
// the near-duplicate filter; bm25_mma.cu: BM25 head-term scoring): PTX wrappers for mbarriers,
// TMA, tcgen05 and TMEM, the admission-bound kernels' launcher, the tensor-map encoder and the
// selection of the KP best fp32 keys out of the CTA-private candidate lists.
#pragma once
#include <cuda.h>

#include "dense_common.cuh"

namespace cmr {

constexpr unsigned long long TMA_EVICT_NORMAL = 0x1000000000000000ull;
constexpr unsigned long long TMA_EVICT_FIRST = 0x12F0000000000000ull;
constexpr unsigned long long TMA_EVICT_LAST = 0x14F0000000000000ull;

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// one non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(u32 bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(u32 dst, const CUtensorMap* map, u32 bar, int c0, int c1,
                                            unsigned long long hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u32 bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(u32 d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                            u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor of a K-major tile with 128-byte swizzle: rows of 64 bf16
// (128 B), groups of 8 rows 1024 B apart (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ unsigned long long umma_desc_sw128(u32 smem_addr) {
  const u32 lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);              // start address, LBO = 1 (unused)
  const u32 hi = (1024u >> 4) | (1u << 14) | (2u << 29);                  // SBO, version, SWIZZLE_128B
  return ((unsigned long long)hi << 32) | lo;
}
// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread, complete on return
__device__ __forceinline__ void tmem_ld32(u32 taddr, float (&v)[32]) {
  u32 r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load split in two: issue (the 32 registers are written asynchronously) and wait (takes the
// registers as in/out operands, so nothing that uses them can be scheduled ahead of it).  Independent
// work placed between the two overlaps the TMEM read latency.
__device__ __forceinline__ void tmem_ld32_issue(u32 taddr, u32 (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(u32 (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// v[j] for a run-time j: registers cannot be indexed, a dense switch becomes one indirect branch
__device__ __forceinline__ float pick32(const float (&v)[32], int j) {
  switch (j) {
#define CMR_PICK(i) case i: return v[i];
    CMR_PICK(0) CMR_PICK(1) CMR_PICK(2) CMR_PICK(3) CMR_PICK(4) CMR_PICK(5) CMR_PICK(6) CMR_PICK(7)
    CMR_PICK(8) CMR_PICK(9) CMR_PICK(10) CMR_PICK(11) CMR_PICK(12) CMR_PICK(13) CMR_PICK(14) CMR_PICK(15)
    CMR_PICK(16) CMR_PICK(17) CMR_PICK(18) CMR_PICK(19) CMR_PICK(20) CMR_PICK(21) CMR_PICK(22) CMR_PICK(23)
    CMR_PICK(24) CMR_PICK(25) CMR_PICK(26) CMR_PICK(27) CMR_PICK(28) CMR_PICK(29) CMR_PICK(30)
#undef CMR_PICK
    default: return v[31];
  }
}


// ---------------------------------------------------------------------------------------
// Candidate lists -> the KP best keys of one query (used by both finalize kernels).
//   cand [n_lists][n_queries][cap] u64 keys (orderable fp32 score, ~row), cnt [n_lists][n_queries]
//   entries appended (may exceed cap: the excess was dropped).
//   s_keys [cap_total], s_out [KP], s_surv [4 * KP]: shared-memory scratch; s_ctl [2] shared ints.
// On return (after a __syncthreads) s_out holds the KP best keys, best first, empties (0) last;
// *n_total = candidates collected (before clamping to cap_total), *overflow = some list or the
// collection buffer overflowed.  Must be called by all FIN_THREADS threads of the CTA.
// ---------------------------------------------------------------------------------------
template <int KP>
__device__ __forceinline__ void cand_collect_select(const u64* __restrict__ cand, const int* __restrict__ cnt,
                                                    int n_lists, int n_queries, int cap, int cap_total, int qi,
                                                    u64* s_keys, u64* s_out, u64* s_surv, int* s_ctl,
                                                    int* n_total, int* overflow) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int& s_total = s_ctl[0];
  int& s_over = s_ctl[1];
  if (tid == 0) {
    s_total = 0;
    s_over = 0;
  }
  for (int i = tid; i < KP; i += FIN_THREADS) s_out[i] = 0ull;
  __syncthreads();
  // Every list's length is read in one round trip (thread l owns list l) and claims its range of the
  // shared buffer; then the warps copy the lists (warp w: lists w, w + 32, ...), all loads independent:
  // two dependent global round trips in all, whatever the number of lists.
  __shared__ int s_base[FIN_THREADS], s_take[FIN_THREADS];
  for (int l0 = 0; l0 < n_lists; l0 += FIN_THREADS) {
    const int l = l0 + tid;
    int have = 0;
    if (l < n_lists) have = cnt[(size_t)l * n_queries + qi];
    const int take = have < cap ? have : cap;
    if (have > cap) s_over = 1;
    s_take[tid] = take;
    s_base[tid] = take > 0 ? atomicAdd(&s_total, take) : 0;
    __syncthreads();
    const int n_here = n_lists - l0 < FIN_THREADS ? n_lists - l0 : FIN_THREADS;
    for (int j = warp; j < n_here; j += FIN_THREADS / 32) {
      const int t2 = s_take[j], b2 = s_base[j];
      const u64* src = cand + ((size_t)(l0 + j) * n_queries + qi) * cap;
      for (int i = lane; i < t2; i += 32) {
        if (b2 + i < cap_total) s_keys[b2 + i] = src[i];
        else s_over = 1;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  *n_total = s_total;
  *overflow = s_over;
  const int m = s_total < cap_total ? s_total : cap_total;
  __syncthreads();  // every thread has read s_total / s_over before s_total is reused as a counter below
  // Ranking m keys against each other is quadratic, and only the KP best matter: find the
  // KP-th largest score word by a radix select (4 passes over one byte each: histogram in shared
  // memory, suffix sums, pick the byte that holds the KP-th), then rank just the keys that reach it
  // (the KP best plus ties of the last one).
  u32 floor_word = 0;
  if (m > 4 * KP) {
    __shared__ int s_hist[256];
    __shared__ u32 s_sel[2];
    u32 prefix = 0;
    int need = KP;
    for (int pass = 3; pass >= 0; --pass) {
      const int shift = pass * 8;
      const u32 hi_mask = pass == 3 ? 0u : (0xFFFFFFFFu << (shift + 8));
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      for (int e = tid; e < m; e += FIN_THREADS) {
        const u32 w = (u32)(s_keys[e] >> 32);
        if ((w & hi_mask) == prefix) atomicAdd(&s_hist[(w >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid < 256) {
        int suf = 0;   // keys (with the decided high bytes) whose byte is >= tid
        for (int x = tid; x < 256; ++x) suf += s_hist[x];
        const int above = suf - s_hist[tid];
        if (suf >= need && above < need) {   // exactly one byte value holds the need-th largest
          s_sel[0] = prefix | ((u32)tid << shift);
          s_sel[1] = (u32)(need - above);
        }
      }
      __syncthreads();
      prefix = s_sel[0];
      need = (int)s_sel[1];
    }
    floor_word = prefix;
  }
  // survivors (score word >= floor) are compacted and ranked among themselves; with more than
  // 4*KP of them (a crowd of ties at the boundary) every key is ranked against all m instead
  if (tid == 0) s_total = 0;
  __syncthreads();
  for (int e = tid; e < m; e += FIN_THREADS) {
    const u64 key = s_keys[e];
    if ((u32)(key >> 32) >= floor_word) {
      const int slot = atomicAdd(&s_total, 1);
      if (slot < 4 * KP) s_surv[slot] = key;
    }
  }
  __syncthreads();
  const int n_surv = s_total;
  if (n_surv <= 4 * KP) {
    for (int e = tid; e < n_surv; e += FIN_THREADS) {
      const u64 key = s_surv[e];
      int rank = 0;
      for (int j = 0; j < n_surv; ++j) rank += s_surv[j] > key;
      if (rank < KP) s_out[rank] = key;
    }
  } else {
    for (int e = tid; e < m; e += FIN_THREADS) {
      const u64 key = s_keys[e];
      if ((u32)(key >> 32) < floor_word) continue;
      int rank = 0;
      for (int j = 0; j < m && rank < KP; ++j) rank += s_keys[j] > key;
      if (rank < KP) s_out[rank] = key;
    }
  }
  __syncthreads();
}

// ---- host helpers defined in dense_mma.cu ------------------------------------------------
// 2-byte element [n_rows, dim] row-major matrix, boxes of [box_rows, 64] columns, 128-byte swizzle,
// out-of-bounds elements read as zero.  f16: element type fp16 instead of bf16 (same bytes moved).
int make_tmap(CUtensorMap* map, const void* ptr, long long n_rows, int dim, int box_rows, bool f16 = false);
// thr[q] = the kp-th largest of gmax[g * gstride + q], g < n_groups (-inf with fewer than kp groups)
int launch_admission_bound(const float* gmax, int n_groups, int gstride, int kp, int n_queries, float* thr,
                           cudaStream_t st);
constexpr int MM_MAX_GROUPS = 49152; // the bound kernel keeps the group maxima in shared memory

}  // namespace cmr
