// Micro-benchmark: legacy mma.sync.m16n8k16 f16 (fp32 accumulate) throughput per SM on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mma_f16 tools/ubench/mma_f16.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024) kern(float* out, int iters, long long* cycles) {
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, threadIdx.x};
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2048;
  for (int threads : {128, 256, 512, 1024}) {
    kern<<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = h[0], inst = double(iters) * 8 * (threads / 32);
    printf("mma.sync m16n8k16 f16, threads/SM %4d: %.0f cycles, %.3f MMA/clk/SM, %.0f FMA/clk/SM\n", threads, c, inst / c, inst * 2048 / c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
