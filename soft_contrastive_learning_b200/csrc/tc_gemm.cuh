// tc_gemm.cuh -- interface of the tcgen05 TF32 / 3xTF32 GEMM (tc_gemm.cu).
#pragma once
#include <cuda_fp16.h>

#include <atomic>

#include "common.cuh"

namespace scl {

// C[b][M,N] (+)= (A[b] . B[b]^T) * rowscale[b][m] * colscale[n]      for b < batch
//   a_mn = false: A(m,k) = A[m*lda + k]   (K-major)      a_mn = true: A(m,k) = A[k*lda + m]   (MN-major)
//   b_mn = false: B(n,k) = B[n*ldb + k]   (K-major)      b_mn = true: B(n,k) = B[k*ldb + n]   (MN-major)
//   precision 0: fp32-grade 3xTF32, 1: one TF32 pass
struct TcGemmDesc {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int lda, ldb, ldc;
  bool a_mn, b_mn;
  const float* colscale;   // optional [N]
  int precision;
  // optional extensions (zero-initialised = off)
  int batch;               // 0/1: single problem; > 1: blockIdx.z, operands at A + b*sA, B + b*sB, C + b*sC
  long long sA, sB, sC;    // batch strides in elements (multiples of 4); sB may be 0 (shared B)
  const float* rowscale;   // optional [batch][M] (+ b*rowscale_stride)
  long long rowscale_stride;
  int accumulate;          // C += result
  int split_k;             // 0/1: off; 2: two CTAs per tile, each half of K, combined with atomicAdd (C zeroed by the caller;
                           //       two addends commute, so the result stays deterministic)
  // optional second operand pair, concatenated along K:  C = [A | A2] . [B | B2]^T  (one pass over C instead of a second,
  // accumulating GEMM).  Same majorness, leading dimensions and batch stride of A as the first pair; K a multiple of 32.
  const float* A2;
  const float* B2;
  int K2;
  long long sB2;           // batch stride of B2 in elements; 0 = shared by every batch
};

struct TcGemmArgs {
  int M, N, K;
  const float* colscale;
  float* C;
  int ldc;
  long long sC;
  const float* rowscale;
  long long rowscale_stride;
  int accumulate;
  int batch;               // tiles = tiles_n * tiles_m * batch * split_k, walked by persistent CTAs
  int split_k;
  int tiles_n, tiles_m;
  int b_shared;            // every batch reads B slice 0
  int k1_stages;           // K stages served by the first operand pair (the rest come from the second pair)
  int b2_shared;
};

int tc_gemm(const TcGemmDesc& d, cudaStream_t stream);

// tc_gemm_h3.cu: fp32-grade GEMM from operands pre-split into fp16 hi / lo halves (kind::f16, three products per k)
//   C[M,N] = rowscale[m] * colscale[n] * (Ah + Al) . (Bh + Bl)^T;  B K-major [N,K] or MN-major [K,N] (b_mn)
int tc_gemm_h3(const __half* Ah, const __half* Al, const __half* Bh, const __half* Bl, float* C, int M, int N, int K, int lda,
               int ldb, int ldc, bool b_mn, const float* rowscale, const float* colscale, int split_k, cudaStream_t stream);
// rows of x (minus `sub`, divided by sqrt(isqrt_of), both optional, per column) -> fp16 hi / lo with one power-of-two scale
// per row; unscale[r] = 2^-e (* *mul)
int h3_split_rows(const float* x, const float* sub, const float* isqrt_of, int R, int Cc, __half* hi, __half* lo,
                  float* unscale, const float* mul, cudaStream_t stream);
// whole matrix with ONE scale (usable as the contraction operand in either orientation); *unscale = 2^-e
int h3_split_all(const float* x, long long n, unsigned int* maxbits, __half* hi, __half* lo, float* unscale,
                 cudaStream_t stream);
int tc_gemm_precision();   // process-wide default set through scl_set_gemm_precision

}  // namespace scl
