// netvlad_fused.cuh -- interface of the fused NetVLAD kernels (netvlad_fused.cu) used by netvlad.cu.
#pragma once
#include "common.cuh"

namespace scl {

// shapes the fused kernels cover (K = 64 clusters, C a multiple of 128 up to 512); false also when SCL_NV_FUSED = 0
bool nv_fused_ok(int B, int HW, int C, int K);
// extra workspace behind the buffers of the generic path: scratch assignment tiles, per-(image, slot) partials
size_t nv_fused_ws_bytes(int B, int HW, int C, int K);
// forward: fills inv [B*HW], a [B*HW,64], V [B,C,64] (un-normalised, centre term included), asum, nk, nt and out [B,C*64]
int nv_fused_fwd(const float* x, const float* assign_w, const float* centers, int B, int HW, int C, float* inv, float* a,
                 float* V, float* asum, float* nk, float* nt, float* out, void* ws, size_t ws_bytes, cudaStream_t stream);

// first half of the backward (da GEMM, soft-max backward, dW) in one pass over x: ds [B*HW,64], rb [B*HW], dW [C,64] (may be
// NULL); same extra workspace as the forward
int nv_fused_bwd(const float* x, const float* a, const float* dV, const float* dasum, int B, int HW, int C, float* ds,
                 float* rb, float* dW, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace scl
