// topk.cuh -- warp-private top-k lists, block merges, cross-CTA selection and
// the exact float64 dot used by the dense rescoring pass.
//
// Everything is generic over a key type K with a strict total order:
//   u64   (dense)  high word = orderable fp32 score, low word = ~row; 0 = empty
//   KeyD  (BM25)   exact float64 score + row; row 0xFFFFFFFF = empty
// "better" means: higher score, then lower row.
#pragma once
#include "common.cuh"

namespace cmr {

struct __align__(16) KeyD {
  double s;
  u32 id;
  u32 pad;
};

__device__ __forceinline__ bool key_gt(u64 a, u64 b) { return a > b; }
__device__ __forceinline__ bool key_empty(u64 a) { return a == 0ull; }
__device__ __forceinline__ void key_clear(u64& a) { a = 0ull; }

__device__ __forceinline__ bool key_gt(const KeyD& a, const KeyD& b) {
  return a.s > b.s || (a.s == b.s && a.id < b.id);
}
__device__ __forceinline__ bool key_empty(const KeyD& a) { return a.id == 0xFFFFFFFFu; }
__device__ __forceinline__ void key_clear(KeyD& a) {
  a.s = -INFINITY;
  a.id = 0xFFFFFFFFu;
  a.pad = 0;
}

__device__ __forceinline__ u64 key_shfl(u64 v, int src) { return shfl_u64(v, src); }
__device__ __forceinline__ u64 key_shfl_up(u64 v, int d) { return shfl_up_u64(v, d); }
__device__ __forceinline__ KeyD key_shfl(const KeyD& v, int src) {
  KeyD r;
  r.s = __shfl_sync(0xFFFFFFFFu, v.s, src);
  r.id = __shfl_sync(0xFFFFFFFFu, v.id, src);
  r.pad = 0;
  return r;
}
__device__ __forceinline__ KeyD key_shfl_up(const KeyD& v, int d) {
  KeyD r;
  r.s = __shfl_up_sync(0xFFFFFFFFu, v.s, d);
  r.id = __shfl_up_sync(0xFFFFFFFFu, v.id, d);
  r.pad = 0;
  return r;
}

// ---------------------------------------------------------------------------
// Warp-private sorted list of KP keys in shared memory (best first, empties
// last).  No lock: only the owning warp touches it.  Insertion is rare once the
// list's last score has warmed up (about KP*(1+ln(rows/KP)) inserts per warp),
// so the streaming loops only pay one score compare against the published
// threshold.  Returns true when the list changed; `new_last` is then its last key.
// ---------------------------------------------------------------------------
template <int KP, typename K>
__device__ __forceinline__ bool warp_list_insert(K* keys, const K& key, int lane, K& new_last) {
  constexpr int KPL = KP / 32;
  K mine[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) mine[j] = keys[j * 32 + lane];
  const K last = key_shfl(mine[KPL - 1], 31);
  bool changed = false;
  if (key_gt(key, last)) {  // warp-uniform and authoritative
    changed = true;
    int pos = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) pos += __popc(__ballot_sync(0xFFFFFFFFu, key_gt(mine[j], key)));
#pragma unroll
    for (int j = KPL - 1; j >= 0; --j) {
      K up = key_shfl_up(mine[j], 1);
      if (j > 0) {
        const K carry = key_shfl(mine[j - 1], 31);
        if (lane == 0) up = carry;
      }
      const int p = j * 32 + lane;
      K nv = mine[j];
      if (p == pos) nv = key;
      else if (p > pos) nv = up;
      keys[p] = nv;
      if (j == KPL - 1) new_last = key_shfl(nv, 31);
    }
  }
  __syncwarp();
  return changed;
}

// number of entries of a best-first sorted list (empties last) that beat k
template <typename K>
__device__ __forceinline__ int count_better(const K* list, int n, const K& k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (key_gt(list[mid], k)) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// number of non-empty entries of a sorted list
template <typename K>
__device__ __forceinline__ int count_valid(const K* list, int n) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (!key_empty(list[mid])) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Merge n_lists sorted lists of KP unique keys (shared memory, list l starts at
// lists + l*stride) into the KP best, written sorted to out[KP] (cleared by the
// caller beforehand).
template <int KP, typename K>
__device__ __forceinline__ void block_merge_lists(const K* lists, int n_lists, size_t stride, K* out,
                                                  int tid, int nthreads) {
  for (int e = tid; e < n_lists * KP; e += nthreads) {
    const int a = e / KP;
    const K k = lists[a * stride + (e - a * KP)];
    if (key_empty(k)) continue;
    int rank = e - a * KP;
    for (int b = 0; b < n_lists && rank < KP; ++b) {
      if (b == a) continue;
      rank += count_better(lists + b * stride, KP, k);
    }
    if (rank < KP) out[rank] = k;
  }
}

constexpr int SELECT_RANK_LISTS = 512;

// Select the KP best keys out of n_lists sorted lists held in GLOBAL memory.
//   s_heads [n_lists + 1] scratch   s_buf [CAP] scratch   s_cnt [1] scratch
//   s_out   [KP] result, sorted best first, empties last
// T = the KP-th best list head is a lower bound of the KP-th best key, and only
// the (at most KP) lists whose head reaches T can hold keys >= T.  Those keys
// (typically KP plus a few) are compacted into s_buf and ranked by counting.
// If more than CAP keys qualify (adversarial data: the bound is KP*KP), each
// qualifying key is instead ranked by binary searches over the qualifying lists
// in global memory -- slower, same result.  Ends with __syncthreads().
template <int KP, int CAP, typename K>
__device__ void block_select_from_lists(const K* __restrict__ lists, int n_lists, K* s_heads, K* s_buf,
                                        int* s_cnt, K* s_out, int tid, int nthreads) {
  for (int i = tid; i < n_lists; i += nthreads) s_heads[i] = lists[(size_t)i * KP];
  for (int i = tid; i < KP; i += nthreads) key_clear(s_out[i]);
  if (tid == 0) {
    s_cnt[0] = 0;
    key_clear(s_heads[n_lists]);  // T: empty == "no bound"
  }
  __syncthreads();
  if (n_lists > KP) {
    // Ranking the heads against each other is quadratic: with very many lists only the first
    // SELECT_RANK_LISTS heads are ranked.  Their KP-th best is still attained by KP distinct
    // keys, so it is still a valid (slightly lower) bound.
    const int n_rank = n_lists < SELECT_RANK_LISTS ? n_lists : SELECT_RANK_LISTS;
    for (int i = tid; i < n_rank; i += nthreads) {
      const K h = s_heads[i];
      if (key_empty(h)) continue;
      int cnt = 0;
      for (int j = 0; j < n_rank; ++j) cnt += key_gt(s_heads[j], h);
      if (cnt == KP - 1) s_heads[n_lists] = h;  // unique keys: exactly one writer
    }
    __syncthreads();
  }
  const K T = s_heads[n_lists];
  const bool bounded = !key_empty(T);
  for (int e = tid; e < n_lists * KP; e += nthreads) {
    const int a = e / KP;
    const K h = s_heads[a];
    if (key_empty(h) || (bounded && key_gt(T, h))) continue;
    const K k = lists[e];
    if (!key_empty(k) && !(bounded && key_gt(T, k))) {
      const int slot = atomicAdd(&s_cnt[0], 1);
      if (slot < CAP) s_buf[slot] = k;
    }
  }
  __syncthreads();
  const int m = s_cnt[0];
  if (m <= CAP) {
    for (int e = tid; e < m; e += nthreads) {
      const K k = s_buf[e];
      int rank = 0;
      for (int j = 0; j < m; ++j) rank += key_gt(s_buf[j], k);
      if (rank < KP) s_out[rank] = k;
    }
  } else {
    for (int e = tid; e < n_lists * KP; e += nthreads) {
      const int a = e / KP;
      const K h = s_heads[a];
      if (key_empty(h) || (bounded && key_gt(T, h))) continue;
      const K k = lists[e];
      if (key_empty(k) || (bounded && key_gt(T, k))) continue;
      int rank = e - a * KP;
      for (int b2 = 0; b2 < n_lists && rank < KP; ++b2) {
        if (b2 == a) continue;
        const K hb = s_heads[b2];
        if (key_empty(hb) || (bounded && key_gt(T, hb))) continue;
        rank += count_better(lists + (size_t)b2 * KP, KP, k);
      }
      if (rank < KP) s_out[rank] = k;
    }
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
