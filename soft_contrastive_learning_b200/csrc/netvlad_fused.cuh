// netvlad_fused.cuh -- interface of the fused NetVLAD kernels (netvlad_fused.cu) used by netvlad.cu.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace scl {

// shapes the fused kernels cover (K = 64 clusters, C a multiple of 128 up to 512); false also when SCL_NV_FUSED = 0
bool nv_fused_ok(int B, int HW, int C, int K);
// extra workspace behind the buffers of the generic path: scratch assignment tiles, per-(image, slot) partials
size_t nv_fused_ws_bytes(int B, int HW, int C, int K);
// forward: fills inv [B*HW], a [B*HW,64], V [B,C,64] (un-normalised, centre term included), asum, nk, nt and out [B,C*64]
int nv_fused_fwd(const float* x, const float* assign_w, const float* centers, int B, int HW, int C, float* inv, float* a,
                 float* V, float* asum, float* nk, float* nt, float* out, void* ws, size_t ws_bytes, cudaStream_t stream);

// the backward in two kernels, one pass over x each: ds [B*HW,64] (d loss / d logits), rb [B*HW] (row term of the
// l2-normalisation backward), dW [C,64] and dx [B,HW,C] (either may be NULL); same extra workspace as the forward, which
// must have been run on it (it leaves the fp16 halves of the soft assignments and of W there)
// head != NULL: dV [B,C,64] and dasum [B,64] are OUTPUTS, computed from the forward's V / nk / nt, d out and the centres by the
// first kernel of the call (the backward of the two normalisations fused with the operand splits); head == NULL: inputs.
struct NvBwdHead {
  const float *V, *dout, *nk, *nt, *centers;
};
int nv_fused_bwd(const float* x, const float* a, const float* inv, float* dV, float* dasum, int B, int HW, int C,
                 float* ds, float* rb, float* dW, float* dx, void* ws, size_t ws_bytes, cudaStream_t stream,
                 const NvBwdHead* head = nullptr);
// netvlad_dx.cu
int nv_dx(const float* x, const __half* a_hi, const __half* a_lo, const __half* ds_hi, const __half* ds_lo, const __half* dv_hi,
          const __half* dv_lo, const __half* w_hi, const __half* w_lo, const float* inv, const float* rb, const float* dvun,
          const float* dsscale, const float* wun, int B, int HW, int C, float* dx, cudaStream_t stream);

}  // namespace scl
