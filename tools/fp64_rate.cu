// fp64_rate.cu -- microbenchmark: vector DFMA vs DMMA (mma.sync m8n8k4 f64) throughput per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/fp64_rate tools/fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

__global__ void cvt_kernel(double* out, const float* in, int iters) {
  float f = in[threadIdx.x];
  double acc = 0;
  long long bits = 0;
  for (int i = 0; i < iters; ++i) {
    // 4 float->double conversions per iteration, kept alive through integer xors
    bits ^= __double_as_longlong((double)f); f += 1.0f;
    bits ^= __double_as_longlong((double)f); f += 1.0f;
    bits ^= __double_as_longlong((double)f); f += 1.0f;
    bits ^= __double_as_longlong((double)f); f += 1.0f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (double)bits;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  float* in;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  cudaMalloc(&in, sizeof(float) * 1024);
  cudaMemset(in, 0, sizeof(float) * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<sms, 1024>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("DFMA: %.3f ms, %.2f lane-FMA/clk/SM (at 1.965 GHz), %.2f TFLOP/s\n", ms,
                    4.0 * iters * 1024 / (ms * 1e-3 * 1.965e9), 2.0 * 4 * iters * 1024.0 * sms / (ms * 1e-3) / 1e12);
    cudaEventRecord(e0);
    dmma_kernel<<<sms, 1024>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("DMMA m8n8k4: %.3f ms, %.2f FMA/clk/SM, %.2f TFLOP/s, %.1f clk per warp-MMA per SM\n", ms,
                    4.0 * iters * 32 * 256 / (ms * 1e-3 * 1.965e9), 2.0 * 4 * iters * 32 * 256.0 * sms / (ms * 1e-3) / 1e12,
                    ms * 1e-3 * 1.965e9 / (4.0 * iters * 32));
    cudaEventRecord(e0);
    cvt_kernel<<<sms, 1024>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("F2F.F64.F32: %.3f ms, %.2f lane-cvt/clk/SM\n", ms, 4.0 * iters * 1024 / (ms * 1e-3 * 1.965e9));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
