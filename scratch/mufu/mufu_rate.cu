// MUFU throughput per function on sm_100a: lanes per clock per SM.  nvcc -arch=sm_100a -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __device__ __forceinline__ float f(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("rsqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) asm volatile("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 4) asm volatile("rcp.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 5) asm volatile("lg2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 6) y = fmaf(x, 1.0001f, 0.5f);
  return y;
}
template <int OP> __global__ void k(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0.5f + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f<OP>(v[i]);
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, float* out, double ghz) {
  const int iters = 4096, blocks = 148 * 2, threads = 1024;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<blocks, threads>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(a); k<OP><<<blocks, threads>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)blocks * threads * iters * 8;
  printf("%-6s %8.3f ms  %6.2f lanes/clk/SM (at %.3f GHz)\n", name, ms, ops / (ms * 1e-3) / (ghz * 1e9) / 148.0, ghz);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 1024 * 4);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6;
  run<0>("tanh", out, ghz); run<1>("rsqrt", out, ghz); run<2>("sqrt", out, ghz); run<3>("ex2", out, ghz);
  run<4>("rcp", out, ghz); run<5>("lg2", out, ghz); run<6>("ffma", out, ghz);
  return 0;
}
