// tc_gemm.cuh -- interface of the tcgen05 TF32 / 3xTF32 GEMM (tc_gemm.cu).
#pragma once
#include <atomic>

#include "common.cuh"

namespace scl {

// C[M,N] = (A . B^T) * colscale[n]
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
};

struct TcGemmArgs {
  int M, N, K;
  const float* colscale;
  float* C;
  int ldc;
};

int tc_gemm(const TcGemmDesc& d, cudaStream_t stream);
int tc_gemm_precision();   // process-wide default set through scl_set_gemm_precision

}  // namespace scl
