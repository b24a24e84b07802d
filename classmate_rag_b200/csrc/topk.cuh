// topk.cuh -- register-resident warp top-k lists, block merges and the exact
// float64 dot used by every rescoring pass.
#pragma once
#include "common.cuh"

namespace cmr {

// A warp keeps its KP = 32*KPL best keys sorted descending; position
// p = j*32 + lane lives in key[j] of that lane.  Insertion is rare once the
// threshold has warmed up, so the scan loop only pays one 64-bit compare.
template <int KPL>
struct WarpList {
  u64 key[KPL];
  u64 kmin;  // key at position KP-1 (warp-uniform); insert only if candidate > kmin

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KPL; ++j) key[j] = 0ull;
    kmin = 0ull;
  }

  // k is warp-uniform and k > kmin.
  __device__ __forceinline__ void insert(u64 k, int lane) {
    int pos = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) pos += __popc(__ballot_sync(0xFFFFFFFFu, key[j] > k));
#pragma unroll
    for (int j = KPL - 1; j >= 0; --j) {
      u64 up = shfl_up_u64(key[j], 1);
      if (j > 0) {
        u64 carry = shfl_u64(key[j - 1], 31);
        if (lane == 0) up = carry;
      }
      const int p = j * 32 + lane;
      if (p == pos) key[j] = k;
      else if (p > pos) key[j] = up;
    }
    kmin = shfl_u64(key[KPL - 1], 31);
  }

  __device__ __forceinline__ void store(u64* dst, int lane) const {
#pragma unroll
    for (int j = 0; j < KPL; ++j) dst[j * 32 + lane] = key[j];
  }
};

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
//   s_heads [n_lists]  scratch      s_stage [KP*KP] scratch
//   s_q     [KP+1]     scratch (int)  s_out [KP] result, sorted descending
// Only lists whose head is among the KP largest heads can contribute, so at
// most KP lists are staged.  Ends with __syncthreads().
template <int KP>
__device__ void block_select_from_lists(const u64* __restrict__ lists, int n_lists,
                                        u64* s_heads, u64* s_stage, int* s_q, u64* s_out,
                                        int tid, int nthreads) {
  for (int i = tid; i < n_lists; i += nthreads) s_heads[i] = lists[(size_t)i * KP];
  for (int i = tid; i < KP; i += nthreads) s_out[i] = 0ull;
  if (tid == 0) s_q[KP] = 0;
  __syncthreads();
  // lists whose head ranks < KP among the heads qualify (keys are unique; empty
  // lists have head 0 and never qualify)
  for (int i = tid; i < n_lists; i += nthreads) {
    const u64 h = s_heads[i];
    if (h == 0ull) continue;
    int cnt = 0;
    for (int j = 0; j < n_lists; ++j) cnt += (s_heads[j] > h);
    if (cnt < KP) {
      int slot = atomicAdd(&s_q[KP], 1);
      s_q[slot] = i;
    }
  }
  __syncthreads();
  const int nq = s_q[KP];
  for (int e = tid; e < nq * KP; e += nthreads) {
    const int a = e / KP;
    s_stage[e] = lists[(size_t)s_q[a] * KP + (e - a * KP)];
  }
  __syncthreads();
  block_merge_lists<KP>(s_stage, nq, s_out, tid, nthreads);
  __syncthreads();
}

// Exact float64 dot of two bf16 vectors in the pinned order of the oracle
// (oracle/np_oracle.py:exact_dots): lane l sums elements l, l+32, ... in order,
// then the halving tree off = 16..1.  All lanes must call; lane 0 holds the
// result (it is also broadcast to every lane).
__device__ __forceinline__ double warp_exact_dot(const uint16_t* __restrict__ a,
                                                 const uint16_t* __restrict__ b, int dim, int lane) {
  double acc = 0.0;
  for (int i = lane; i < dim; i += 32) {
    // the product of two bf16 values is exact in float64, so fma == mul then add
    acc = __fma_rn(bf16_to_f64(a[i]), bf16_to_f64(b[i]), acc);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    double other = __shfl_down_sync(0xFFFFFFFFu, acc, off);
    acc = __dadd_rn(acc, other);
  }
  return __shfl_sync(0xFFFFFFFFu, acc, 0);
}

}  // namespace cmr
