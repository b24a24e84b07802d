// topk.cuh -- CTA-shared top-k lists, block merges and the exact float64 dot
// used by every rescoring pass.
#pragma once
#include "common.cuh"

namespace cmr {

// ---------------------------------------------------------------------------
// Warp-private sorted list of KP keys in shared memory (descending, 0 = empty)
// with the fp32 score of its last entry published as the warp's admission
// threshold.  No lock: only the owning warp touches it.  Insertion is rare once
// the threshold has warmed up (about KP*(1+ln(rows/KP)) inserts per warp), so
// the streaming loops only pay one float compare per score against `thr`.
// ---------------------------------------------------------------------------
template <int KP>
__device__ __forceinline__ void warp_list_insert(u64* keys, float* thr, u64 key, int lane) {
  constexpr int KPL = KP / 32;
  u64 mine[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) mine[j] = keys[j * 32 + lane];
  const u64 last = shfl_u64(mine[KPL - 1], 31);
  if (key > last) {  // warp-uniform, authoritative (the float pre-test admits ties)
    int pos = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) pos += __popc(__ballot_sync(0xFFFFFFFFu, mine[j] > key));
#pragma unroll
    for (int j = KPL - 1; j >= 0; --j) {
      u64 up = shfl_up_u64(mine[j], 1);
      if (j > 0) {
        const u64 carry = shfl_u64(mine[j - 1], 31);
        if (lane == 0) up = carry;
      }
      const int p = j * 32 + lane;
      u64 nv = mine[j];
      if (p == pos) nv = key;
      else if (p > pos) nv = up;
      keys[p] = nv;
      if (j == KPL - 1 && lane == 31 && nv != 0ull) *thr = key_score(nv);
    }
  }
  __syncwarp();
}

// number of entries of a descending-sorted list (zeros at the end) that are > k
__device__ __forceinline__ int count_greater(const u64* list, int n, u64 k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (list[mid] > k) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Merge n_lists descending-sorted lists of KP unique keys (shared memory) into
// the KP best, written sorted to out[KP] (must be zero-filled by the caller).
template <int KP>
__device__ __forceinline__ void block_merge_lists(const u64* lists, int n_lists, u64* out,
                                                  int tid, int nthreads) {
  for (int e = tid; e < n_lists * KP; e += nthreads) {
    const u64 k = lists[e];
    if (k == 0ull) continue;
    const int a = e / KP;
    int rank = e - a * KP;
    for (int b = 0; b < n_lists && rank < KP; ++b) {
      if (b == a) continue;
      rank += count_greater(lists + b * KP, KP, k);
    }
    if (rank < KP) out[rank] = k;
  }
}

// Select the KP best keys out of n_lists sorted lists held in GLOBAL memory.
//   s_heads [n_lists] scratch   s_buf [KP*KP] scratch   s_cnt [2] scratch (int)
//   s_out   [KP] result, sorted descending, zero padded
// T = the KP-th largest list head is a lower bound of the KP-th best key, and
// only the (at most KP) lists whose head reaches T can hold keys >= T, so at
// most KP*KP keys are compacted; they are then ranked by counting.  Ends with
// __syncthreads().
template <int KP>
__device__ void block_select_from_lists(const u64* __restrict__ lists, int n_lists,
                                        u64* s_heads, u64* s_buf, int* s_cnt, u64* s_out,
                                        int tid, int nthreads) {
  for (int i = tid; i < n_lists; i += nthreads) s_heads[i] = lists[(size_t)i * KP];
  for (int i = tid; i < KP; i += nthreads) s_out[i] = 0ull;
  if (tid == 0) {
    s_cnt[0] = 0;
    s_heads[n_lists] = 0ull;  // T
  }
  __syncthreads();
  if (n_lists > KP) {
    for (int i = tid; i < n_lists; i += nthreads) {
      const u64 h = s_heads[i];
      if (h == 0ull) continue;
      int cnt = 0;
      for (int j = 0; j < n_lists; ++j) cnt += (s_heads[j] > h);
      if (cnt == KP - 1) s_heads[n_lists] = h;  // unique keys: exactly one writer
    }
    __syncthreads();
  }
  const u64 T = s_heads[n_lists];
  // compact keys >= T (a prefix of each qualifying list); one thread per list slot
  for (int e = tid; e < n_lists * KP; e += nthreads) {
    const int a = e / KP;
    if (s_heads[a] < T || s_heads[a] == 0ull) continue;
    const u64 k = lists[e];
    if (k != 0ull && k >= T) s_buf[atomicAdd(&s_cnt[0], 1)] = k;
  }
  __syncthreads();
  const int m = s_cnt[0];
  for (int e = tid; e < m; e += nthreads) {
    const u64 k = s_buf[e];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += (s_buf[j] > k);
    if (rank < KP) s_out[rank] = k;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Exact float64 dot of two bf16 vectors in the pinned order of the oracle
// (oracle/np_oracle.py:exact_dots): the row is cut into 16-byte vectors of 8
// elements, lane l owns vectors l, l+32, ... and adds their exact products in
// element order; then the halving tree off = 16..1.  a and b must be 16-byte
// aligned, dim % 8 == 0.  All lanes call; every lane gets the result.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fma_word(u32 wa, u32 wb, double acc) {
  acc = __fma_rn((double)bf16lo(wa), (double)bf16lo(wb), acc);
  acc = __fma_rn((double)bf16hi(wa), (double)bf16hi(wb), acc);
  return acc;
}

__device__ __forceinline__ double warp_exact_dot(const uint16_t* __restrict__ a,
                                                 const uint16_t* __restrict__ b, int dim, int lane) {
  const uint4* va = reinterpret_cast<const uint4*>(a);
  const uint4* vb = reinterpret_cast<const uint4*>(b);
  const int nvec = dim >> 3;
  double acc = 0.0;
  // the product of two bf16 values is exact in float64, so fma == mul then add
  for (int v = lane; v < nvec; v += 32) {
    const uint4 x = va[v];
    const uint4 y = vb[v];
    acc = fma_word(x.x, y.x, acc);
    acc = fma_word(x.y, y.y, acc);
    acc = fma_word(x.z, y.z, acc);
    acc = fma_word(x.w, y.w, acc);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double other = __shfl_down_sync(0xFFFFFFFFu, acc, off);
    acc = __dadd_rn(acc, other);
  }
  return __shfl_sync(0xFFFFFFFFu, acc, 0);
}

}  // namespace cmr
