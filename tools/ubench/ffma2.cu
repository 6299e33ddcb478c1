// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue/throughput per SM on sm_100a, and LDS.128 broadcast cost.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma2 tools/ubench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int MODE>
__global__ void __launch_bounds__(1024) kern(float* out, int iters, long long* cycles) {
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  unsigned long long a2, b2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(a), "f"(a + 1.f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(b), "f"(b));
  float acc[16];
  unsigned long long acc2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = i; acc2[i] = i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a), "f"(b));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) ffma2(acc2[i], a2, b2);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i] + float(acc2[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      if (mode == 0) kern<0><<<148, threads>>>(out, iters, cyc); else kern<1><<<148, threads>>>(out, iters, cyc);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = h[0];
      double inst = double(iters) * 16 * (threads / 32);          // warp instructions per SM
      double fma = inst * 32 * (mode ? 2 : 1);
      printf("%s threads/SM %4d: %.0f cycles, %.3f warp-inst/clk/SM, %.1f FMA/clk/SM\n", mode ? "FFMA2" : "FFMA ", threads, c,
             inst / c, fma / c);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
