// sgemm.cuh -- FP32 (FFMA) tiled GEMM used where the reference needs fp32-grade dot products and the
// contraction is not yet on the tcgen05 path: flat-mode Gram / backward, PCA projection, NetVLAD pieces.
//
//   C[b] = epilogue( op(A[b]) * op(B[b]) ),   op selected by transA / transB, optional per-k transform on A
//
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered shared memory.
#pragma once
#include "common.cuh"

namespace scl {

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int lda, ldb, ldc;
  long long sA, sB, sC;    // batch strides (elements)
  int batch;
  int transA;              // 0: A(m,k) = A[m*lda+k]   1: A(m,k) = A[k*lda+m]
  int transB;              // 0: B(k,n) = B[k*ldb+n]   1: B(k,n) = B[n*ldb+k]
  const float* a_sub_k;    // optional: A(m,k) -= a_sub_k[k]
  const float* a_mul_k;    // optional: A(m,k) *= a_mul_k[k]   (applied after the subtraction; + batch*a_mul_k_stride)
  long long a_mul_k_stride;
  const float* a_isqrt_k;  // optional: A(m,k) /= sqrt(a_isqrt_k[k])
  const float* row_scale;  // optional epilogue: acc *= row_scale[m]   (+ batch*row_scale_stride)
  long long row_scale_stride;
  const float* col_scale;  // optional epilogue: acc *= col_scale[n]
  const float* col_isqrt;  // optional epilogue: acc /= sqrt(col_isqrt[n])
  float alpha;
  int accumulate;          // C += result instead of C = result
  int split_k;             // >1: gridDim.z = batch*split_k, partial results atomically added into C (C must be zeroed)
};

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 16, kGemmThreads = 256;

__device__ __forceinline__ float gemm_load_a(const GemmArgs& g, const float* A, const float* amul, int m, int k) {
  if (m >= g.M || k >= g.K) return 0.0f;
  float v = g.transA ? A[size_t(k) * g.lda + m] : A[size_t(m) * g.lda + k];
  if (g.a_sub_k) v -= g.a_sub_k[k];
  if (amul) v *= amul[k];
  if (g.a_isqrt_k) v /= sqrtf(g.a_isqrt_k[k]);
  return v;
}
__device__ __forceinline__ float gemm_load_b(const GemmArgs& g, const float* B, int k, int n) {
  if (n >= g.N || k >= g.K) return 0.0f;
  return g.transB ? B[size_t(n) * g.ldb + k] : B[size_t(k) * g.ldb + n];
}

static __global__ void __launch_bounds__(kGemmThreads) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][kGemmBK][kGemmBM + 4];
  __shared__ __align__(16) float Bs[2][kGemmBK][kGemmBN + 4];
  const int zb = blockIdx.z / g.split_k, zk = blockIdx.z - zb * g.split_k;
  const float* A = g.A + size_t(zb) * g.sA;
  const float* B = g.B + size_t(zb) * g.sB;
  float* C = g.C + size_t(zb) * g.sC;
  const float* amul = g.a_mul_k ? g.a_mul_k + size_t(zb) * g.a_mul_k_stride : nullptr;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;     // 16 x 16 threads, each 8 x 8 outputs (strided by 16 in 4-wide groups)

  // K range of this split
  const int kchunks = (g.K + kGemmBK - 1) / kGemmBK;
  const int per = (kchunks + g.split_k - 1) / g.split_k;
  const int kc_begin = zk * per, kc_end = min(kchunks, kc_begin + per);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  // global -> register staging: each thread moves 8 A values and 8 B values per K chunk
  float ra[8], rb[8];
  auto fetch = [&](int kc) {
    const int k0 = kc * kGemmBK;
    if (g.transA) {   // contiguous along m: thread -> (k = tid/16, m = (tid%16)*8 .. +7)
      const int k = tid >> 4, mb = (tid & 15) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) ra[u] = gemm_load_a(g, A, amul, m0 + mb + u, k0 + k);
    } else {          // contiguous along k: thread -> (m = tid/2, k = (tid%2)*8 .. +7)
      const int m = tid >> 1, kb = (tid & 1) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) ra[u] = gemm_load_a(g, A, amul, m0 + m, k0 + kb + u);
    }
    if (g.transB) {   // B stored [n][k]: contiguous along k
      const int n = tid >> 1, kb = (tid & 1) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) rb[u] = gemm_load_b(g, B, k0 + kb + u, n0 + n);
    } else {          // B stored [k][n]: contiguous along n
      const int k = tid >> 4, nb = (tid & 15) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) rb[u] = gemm_load_b(g, B, k0 + k, n0 + nb + u);
    }
  };
  auto stash = [&](int buf) {
    if (g.transA) {
      const int k = tid >> 4, mb = (tid & 15) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) As[buf][k][mb + u] = ra[u];
    } else {
      const int m = tid >> 1, kb = (tid & 1) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) As[buf][kb + u][m] = ra[u];
    }
    if (g.transB) {
      const int n = tid >> 1, kb = (tid & 1) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) Bs[buf][kb + u][n] = rb[u];
    } else {
      const int k = tid >> 4, nb = (tid & 15) * 8;
#pragma unroll
      for (int u = 0; u < 8; ++u) Bs[buf][k][nb + u] = rb[u];
    }
  };

  if (kc_begin < kc_end) {
    fetch(kc_begin);
    stash(0);
    __syncthreads();
    for (int kc = kc_begin; kc < kc_end; ++kc) {
      const int buf = (kc - kc_begin) & 1;
      if (kc + 1 < kc_end) fetch(kc + 1);
#pragma unroll
      for (int k = 0; k < kGemmBK; ++k) {
        float a[8], b[8];
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (kc + 1 < kc_end) {
        stash(buf ^ 1);
        __syncthreads();
      }
    }
  }

  const float* rs = g.row_scale ? g.row_scale + size_t(zb) * g.row_scale_stride : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    const float rsc = rs ? rs[m] : 1.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha * rsc;
      if (g.col_scale) v *= g.col_scale[n];
      if (g.col_isqrt) v /= sqrtf(g.col_isqrt[n]);
      float* dst = C + size_t(m) * g.ldc + n;
      if (g.split_k > 1) atomicAdd(dst, v);
      else if (g.accumulate) *dst += v;
      else *dst = v;
    }
  }
}

static inline GemmArgs gemm_args(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                                 int transA, int transB) {
  GemmArgs g = {};
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = K;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.batch = 1;
  g.transA = transA; g.transB = transB;
  g.alpha = 1.0f;
  g.split_k = 1;
  return g;
}

static inline int gemm_launch(const GemmArgs& g, cudaStream_t stream) {
  dim3 grid((g.N + kGemmBN - 1) / kGemmBN, (g.M + kGemmBM - 1) / kGemmBM, g.batch * g.split_k);
  sgemm_kernel<<<grid, kGemmThreads, 0, stream>>>(g);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

}  // namespace scl
