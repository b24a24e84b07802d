// common.cuh -- shared device/host helpers for libcmrag (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cmrag.h"

namespace cmr {

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- error plumbing --------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();          // cached, current device
int fail_cuda(cudaError_t e, const char* what);

#define CMR_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      cmr::set_error(__VA_ARGS__);          \
      return CMR_EINVAL;                    \
    }                                       \
  } while (0)

#define CMR_CUDA(call)                                        \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return cmr::fail_cuda(_e, #call);  \
  } while (0)

// ---- ordered keys ----------------------------------------------------------
// A candidate is one 64-bit key: high word = fp32 score mapped to an unsigned
// integer that sorts the same way, low word = ~local_row.  Larger key == better
// (higher score, then lower row).  0 is the empty sentinel.
__device__ __forceinline__ u32 f32_orderable(float f) {
  u32 u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float orderable_f32(u32 o) {
  u32 u = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
  return __uint_as_float(u);
}
__device__ __forceinline__ u64 make_key(float score, u32 row) {
  return ((u64)f32_orderable(score) << 32) | (u64)(0xFFFFFFFFu - row);
}
__device__ __forceinline__ u32 key_row(u64 key) { return 0xFFFFFFFFu - (u32)(key & 0xFFFFFFFFu); }
__device__ __forceinline__ float key_score(u64 key) { return orderable_f32((u32)(key >> 32)); }

__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
  u32 lo = __shfl_sync(0xFFFFFFFFu, (u32)v, src);
  u32 hi = __shfl_sync(0xFFFFFFFFu, (u32)(v >> 32), src);
  return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_up_u64(u64 v, int d) {
  u32 lo = __shfl_up_sync(0xFFFFFFFFu, (u32)v, d);
  u32 hi = __shfl_up_sync(0xFFFFFFFFu, (u32)(v >> 32), d);
  return ((u64)hi << 32) | lo;
}

// streaming 16-byte load: read-only path, do not allocate in L1
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float bf16lo(u32 u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(u32 u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ double bf16_to_f64(uint16_t b) { return (double)__uint_as_float(((u32)b) << 16); }

}  // namespace cmr
