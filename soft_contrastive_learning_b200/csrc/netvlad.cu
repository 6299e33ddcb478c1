// netvlad.cu -- N1 NetVLAD aggregation head (forward + backward) and P1 PCA-whitening projection.
//
// N1 replaces  x = tf.nn.l2_normalize(x, axis=-1); x = layers.netVLAD(x, 64)   (/root/reference/model/nets.py:66-67)
//   xh = x / max(|x|, 1e-6)                                   per spatial position
//   a  = softmax_k(xh W)                                      soft assignment, W = 'assignment/kernel' [C,K]
//   V[c,k] = sum_n a[n,k] (xh[n,c] + Cc[c,k])                 Cc = 'cluster_centers' (stored negated upstream)
//   V[:,k] /= sqrt(sum_c V[c,k]^2 + 1e-12)                    intra-normalisation
//   out = flatten_{c*K+k}(V) / sqrt(sum V^2 + 1e-12)
// The [B,h,w,C,K] residual tensor of the upstream graph is never formed: V = X^T A + Cc * colsum(A).
// P1 replaces train/train.py:650-651:  y = ((x - m) V^T) / sqrt(var).
//
// All six contractions (assignment logits, aggregation, and the four of the backward) run on the tcgen05 GEMM of
// tc_gemm.cu (fp32-grade 3xTF32 by default): x is read K-major for the logits and MN-major -- the same buffer, no
// transposed copy -- for the aggregation; the l2-normalisation of x is never materialised (row scale in the epilogue,
// or folded into the soft assignments).  The row-normalisation, soft-max and the two vector norms are fused
// warp-shuffle kernels.
#include <algorithm>

#include <cuda_fp16.h>

#include "netvlad_fused.cuh"
#include "sgemm.cuh"
#include "tc_gemm.cuh"

namespace scl {

struct NvWs {
  float* inv;     // [B*HW]   1/|x|
  float* a;       // [B*HW,K] soft assignments
  float* V;       // [B,C,K]  un-normalised VLAD (after the centre term)
  float* asum;    // [B,K]
  float* nk;      // [B,K]    intra norms
  float* nt;      // [B]      total norms
  float* dV;      // [B,C,K]
  float* da;      // [B*HW,K] (backward scratch: da then ds)
  float* dasum;   // [B,K]
  float* part;    // [B,C,K] per-image partial dW (backward)
  float* rb;      // [B*HW]  row term of the l2-normalisation backward (fused backward)
};

static size_t nv_ws_bytes(int B, int HW, int C, int K) {
  size_t n = 0;
  n += carve_bytes(size_t(B) * HW, 4);
  n += 2 * carve_bytes(size_t(B) * HW * K, 4);
  n += 3 * carve_bytes(size_t(B) * C * K, 4);
  n += 3 * carve_bytes(size_t(B) * K, 4);
  n += carve_bytes(B, 4);
  n += carve_bytes(size_t(B) * HW, 4);
  return n;
}
// the fused kernels' scratch sits behind the buffers both paths share
static size_t nv_ws_total(int B, int HW, int C, int K) {
  return nv_ws_bytes(B, HW, C, K) + (nv_fused_ok(B, HW, C, K) ? nv_fused_ws_bytes(B, HW, C, K) : 0);
}
static NvWs nv_carve(void* p, size_t bytes, int B, int HW, int C, int K) {
  Carver c(p, bytes);
  NvWs w;
  w.inv = c.take<float>(size_t(B) * HW);
  w.a = c.take<float>(size_t(B) * HW * K);
  w.da = c.take<float>(size_t(B) * HW * K);
  w.V = c.take<float>(size_t(B) * C * K);
  w.dV = c.take<float>(size_t(B) * C * K);
  w.part = c.take<float>(size_t(B) * C * K);
  w.asum = c.take<float>(size_t(B) * K);
  w.nk = c.take<float>(size_t(B) * K);
  w.dasum = c.take<float>(size_t(B) * K);
  w.nt = c.take<float>(B);
  w.rb = c.take<float>(size_t(B) * HW);
  return w;
}

// inv[p] = rsqrt(max(sum_c x[p,c]^2, 1e-12))  -- tf.nn.l2_normalize (nets.py:66).  One warp per position.
__global__ void __launch_bounds__(256) nv_rownorm_kernel(const float* __restrict__ x, long long P, int C,
                                                         float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const float4* row = reinterpret_cast<const float4*>(x + size_t(p) * C);
  float s = 0.0f;
  for (int c = lane; c < (C >> 2); c += 32) {
    const float4 v = __ldg(row + c);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  s = warp_sum(s);
  if (lane == 0) inv[p] = rsqrtf(fmaxf(s, 1e-12f));
}

// in-place softmax over K = 64 columns; one warp per position (2 values per lane).  a_scaled = a * inv[p] is the
// operand of the aggregation GEMM (the l2-normalisation of x folded into the assignments).
__global__ void __launch_bounds__(256) nv_softmax_kernel(float* __restrict__ a, long long P, const float* __restrict__ inv,
                                                         float* __restrict__ a_scaled) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  float* row = a + size_t(p) * 64;
  float v0 = row[lane], v1 = row[lane + 32];
  const float m = warp_max(fmaxf(v0, v1));
  v0 = expf(v0 - m);
  v1 = expf(v1 - m);
  const float s = warp_sum(v0 + v1);
  v0 /= s;
  v1 /= s;
  row[lane] = v0;
  row[lane + 32] = v1;
  const float iv = inv[p];
  a_scaled[size_t(p) * 64 + lane] = v0 * iv;
  a_scaled[size_t(p) * 64 + lane + 32] = v1 * iv;
}

// rows of a [P,64] matrix scaled in place by inv[p]
__global__ void __launch_bounds__(256) nv_scale_rows_kernel(float* __restrict__ a, long long P, const float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const float iv = inv[p];
  a[size_t(p) * 64 + lane] *= iv;
  a[size_t(p) * 64 + lane + 32] *= iv;
}

// out[i] = sum_b part[b][i]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) nv_sum_batch_kernel(const float* __restrict__ part, int B, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.0f;
  for (int b = 0; b < B; ++b) acc += part[size_t(b) * n + i];
  out[i] = acc;
}

// asum[b,k] = sum_n a[b,n,k].  One CTA (256 threads) per image: 4 row groups x 64 columns.
__global__ void __launch_bounds__(256) nv_colsum_kernel(const float* __restrict__ a, int HW, float* __restrict__ asum) {
  __shared__ float s[4][64];
  const int b = blockIdx.x, k = threadIdx.x & 63, g = threadIdx.x >> 6;
  const float* A = a + size_t(b) * HW * 64;
  float acc = 0.0f;
  for (int n = g; n < HW; n += 4) acc += A[size_t(n) * 64 + k];
  s[g][k] = acc;
  __syncthreads();
  if (g == 0) asum[b * 64 + k] = (s[0][k] + s[1][k]) + (s[2][k] + s[3][k]);
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.0f;
  for (int w = 0; w < 8; ++w) r += sh[w];
  return r;
}

// V += Cc * asum; intra-norm per cluster; flatten; l2 norm.  One CTA per image, K = 64.
__global__ void __launch_bounds__(256) nv_norm_fwd_kernel(float* __restrict__ V, const float* __restrict__ centers,
                                                          const float* __restrict__ asum, int C, float* __restrict__ nk,
                                                          float* __restrict__ nt, float* __restrict__ out) {
  __shared__ float s_col[4][64];
  __shared__ float s_nk[64];
  __shared__ float s_red[8];
  const int b = blockIdx.x, k = threadIdx.x & 63, g = threadIdx.x >> 6;
  float* Vb = V + size_t(b) * C * 64;
  const float as = asum[b * 64 + k];
  float ss = 0.0f;
  for (int c = g; c < C; c += 4) {
    const float v = Vb[size_t(c) * 64 + k] + centers[size_t(c) * 64 + k] * as;
    Vb[size_t(c) * 64 + k] = v;
    ss = fmaf(v, v, ss);
  }
  s_col[g][k] = ss;
  __syncthreads();
  if (g == 0) {
    const float n = sqrtf((s_col[0][k] + s_col[1][k]) + (s_col[2][k] + s_col[3][k]) + 1e-12f);
    s_nk[k] = n;
    nk[b * 64 + k] = n;
  }
  __syncthreads();
  const float ink = 1.0f / s_nk[k];
  float tot = 0.0f;
  for (int c = g; c < C; c += 4) {
    const float v = Vb[size_t(c) * 64 + k] * ink;
    tot = fmaf(v, v, tot);
  }
  tot = block_sum_256(tot, s_red);
  const float n_t = sqrtf(tot + 1e-12f);
  if (threadIdx.x == 0) nt[b] = n_t;
  const float sc = ink / n_t;
  float* ob = out + size_t(b) * C * 64;
  for (int c = g; c < C; c += 4) ob[size_t(c) * 64 + k] = Vb[size_t(c) * 64 + k] * sc;
}

// backward of the two norms: dV from dout.  One CTA per image.
//   out = V1 / nt, V1 = V / nk:   dV1 = (dout - out (out.dout)) / nt ;  dV = (dV1 - V1 (V1.dV1)_c) / nk
__global__ void __launch_bounds__(256) nv_norm_bwd_kernel(const float* __restrict__ V, const float* __restrict__ dout,
                                                          const float* __restrict__ nk, const float* __restrict__ nt,
                                                          int C, float* __restrict__ dV) {
  __shared__ float s_col[4][64];
  __shared__ float s_dot[64];
  __shared__ float s_red[8];
  const int b = blockIdx.x, k = threadIdx.x & 63, g = threadIdx.x >> 6;
  const float* Vb = V + size_t(b) * C * 64;
  const float* db = dout + size_t(b) * C * 64;
  float* dvb = dV + size_t(b) * C * 64;
  const float ink = 1.0f / nk[b * 64 + k], intt = 1.0f / nt[b];
  float dot = 0.0f;
  for (int c = g; c < C; c += 4) dot = fmaf(Vb[size_t(c) * 64 + k] * ink * intt, db[size_t(c) * 64 + k], dot);
  dot = block_sum_256(dot, s_red);                 // out . dout
  float cd = 0.0f;
  for (int c = g; c < C; c += 4) {
    const float v1 = Vb[size_t(c) * 64 + k] * ink;
    const float dv1 = (db[size_t(c) * 64 + k] - v1 * intt * dot) * intt;
    cd = fmaf(v1, dv1, cd);
  }
  s_col[g][k] = cd;
  __syncthreads();
  if (g == 0) s_dot[k] = (s_col[0][k] + s_col[1][k]) + (s_col[2][k] + s_col[3][k]);
  __syncthreads();
  const float cdk = s_dot[k];
  for (int c = g; c < C; c += 4) {
    const float v1 = Vb[size_t(c) * 64 + k] * ink;
    const float dv1 = (db[size_t(c) * 64 + k] - v1 * intt * dot) * intt;
    dvb[size_t(c) * 64 + k] = (dv1 - v1 * cdk) * ink;
  }
}

// dcenters[c,k] = sum_b dV[b,c,k] asum[b,k]
__global__ void __launch_bounds__(256) nv_dcenters_kernel(const float* __restrict__ dV, const float* __restrict__ asum,
                                                          int B, int C, float* __restrict__ dcenters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * 64) return;
  const int k = i & 63;
  // four independent chains (the 256 dependent loads of one chain made this 27 us at config 2); fixed order
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  int b = 0;
  for (; b + 4 <= B; b += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = fmaf(dV[size_t(b + u) * C * 64 + i], asum[(b + u) * 64 + k], acc[u]);
  }
  for (; b < B; ++b) acc[0] = fmaf(dV[size_t(b) * C * 64 + i], asum[b * 64 + k], acc[0]);
  dcenters[i] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// dasum[b,k] = sum_c dV[b,c,k] Cc[c,k]
__global__ void __launch_bounds__(256) nv_dasum_kernel(const float* __restrict__ dV, const float* __restrict__ centers,
                                                       int C, float* __restrict__ dasum) {
  __shared__ float s[4][64];
  const int b = blockIdx.x, k = threadIdx.x & 63, g = threadIdx.x >> 6;
  const float* dvb = dV + size_t(b) * C * 64;
  float acc = 0.0f;
  for (int c = g; c < C; c += 4) acc = fmaf(dvb[size_t(c) * 64 + k], centers[size_t(c) * 64 + k], acc);
  s[g][k] = acc;
  __syncthreads();
  if (g == 0) dasum[b * 64 + k] = (s[0][k] + s[1][k]) + (s[2][k] + s[3][k]);
}

// ds = a * ((da + dasum) - sum_k a (da + dasum)); in place on da.  One warp per position.
__global__ void __launch_bounds__(256) nv_softmax_bwd_kernel(const float* __restrict__ a, float* __restrict__ da,
                                                             const float* __restrict__ dasum, long long P, int HW) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const int b = int(p / HW);
  const float a0 = a[size_t(p) * 64 + lane], a1 = a[size_t(p) * 64 + lane + 32];
  const float g0 = da[size_t(p) * 64 + lane] + dasum[b * 64 + lane];
  const float g1 = da[size_t(p) * 64 + lane + 32] + dasum[b * 64 + lane + 32];
  const float dot = warp_sum(a0 * g0 + a1 * g1);
  da[size_t(p) * 64 + lane] = a0 * (g0 - dot);
  da[size_t(p) * 64 + lane + 32] = a1 * (g1 - dot);
}

// dx = inv * (dxh - xh (xh . dxh)), xh = x * inv; in place on dx (which holds dxh).  One warp per position.
__global__ void __launch_bounds__(256) nv_l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ inv,
                                                            long long P, int C, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const float iv = inv[p];
  const bool clamped = iv >= 1e6f * 0.999f;       // sum x^2 below 1e-12: the normalisation is a pure scale
  const float4* xr = reinterpret_cast<const float4*>(x + size_t(p) * C);
  float4* dr = reinterpret_cast<float4*>(dx + size_t(p) * C);
  float dot = 0.0f;
  for (int c = lane; c < (C >> 2); c += 32) {
    const float4 xv = __ldg(xr + c), dv = dr[c];
    dot = fmaf(xv.x * iv, dv.x, dot); dot = fmaf(xv.y * iv, dv.y, dot);
    dot = fmaf(xv.z * iv, dv.z, dot); dot = fmaf(xv.w * iv, dv.w, dot);
  }
  dot = warp_sum(dot);
  if (clamped) dot = 0.0f;
  for (int c = lane; c < (C >> 2); c += 32) {
    const float4 xv = __ldg(xr + c);
    float4 dv = dr[c];
    dv.x = iv * (dv.x - xv.x * iv * dot); dv.y = iv * (dv.y - xv.y * iv * dot);
    dv.z = iv * (dv.z - xv.z * iv * dot); dv.w = iv * (dv.w - xv.w * iv * dot);
    dr[c] = dv;
  }
}

static int nv_check(int B, int HW, int C, int K) {
  if (B < 1 || HW < 1 || C < 4 || (C & 3)) return SCL_ERR_BAD_SHAPE;
  if (K != 64) return SCL_ERR_UNSUPPORTED;     // the reference only ever calls netVLAD(x, 64)
  return SCL_OK;
}

}  // namespace scl

using namespace scl;

extern "C" int scl_netvlad_workspace_bytes(int B, int HW, int C, int K, size_t* bytes) {
  if (!bytes) return SCL_ERR_BAD_ARG;
  int rc = nv_check(B, HW, C, K);
  if (rc) return rc;
  *bytes = nv_ws_total(B, HW, C, K);
  return SCL_OK;
}

extern "C" int scl_netvlad_fwd(const float* x, const float* assign_w, const float* centers, int B, int HW, int C, int K,
                               float* out, void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!x || !assign_w || !centers || !out || !workspace) return SCL_ERR_BAD_ARG;
  int rc = nv_check(B, HW, C, K);
  if (rc) return rc;
  if (!aligned16(x) || !aligned16(out) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return SCL_ERR_ALIGN;
  if (workspace_bytes < nv_ws_total(B, HW, C, K)) return SCL_ERR_WORKSPACE;
  rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NvWs w = nv_carve(workspace, workspace_bytes, B, HW, C, K);
  const long long P = (long long)B * HW;
  if (nv_fused_ok(B, HW, C, K) && tc_gemm_precision() == 0) {
    // one pass over x: netvlad_fused.cu (the backward finds inv, a, V, asum, nk, nt where the generic path leaves them)
    const size_t base = nv_ws_bytes(B, HW, C, K);
    rc = nv_fused_fwd(x, assign_w, centers, B, HW, C, w.inv, w.a, w.V, w.asum, w.nk, w.nt, out,
                      static_cast<char*>(workspace) + base, workspace_bytes - base, stream);
    if (rc != SCL_ERR_UNSUPPORTED) return rc;
  }
  nv_rownorm_kernel<<<unsigned((P + 7) / 8), 256, 0, stream>>>(x, P, C, w.inv);
  SCL_LAUNCH_CHECK();
  const int prec = tc_gemm_precision();
  // logits = (X W) * inv[row]: X K-major, W [C,K] read MN-major
  {
    TcGemmDesc d = {};
    d.A = x; d.B = assign_w; d.C = w.a; d.M = int(P); d.N = K; d.K = C; d.lda = C; d.ldb = K; d.ldc = K;
    d.a_mn = false; d.b_mn = true; d.rowscale = w.inv; d.precision = prec;
    rc = tc_gemm(d, stream);
    if (rc) return rc;
  }
  nv_softmax_kernel<<<unsigned((P + 7) / 8), 256, 0, stream>>>(w.a, P, w.inv, w.da);      // w.da: a * inv (scratch)
  SCL_LAUNCH_CHECK();
  nv_colsum_kernel<<<B, 256, 0, stream>>>(w.a, HW, w.asum);
  SCL_LAUNCH_CHECK();
  // V[b] = X[b]^T (A[b] * inv)     M = C, N = K, contraction over the HW positions; both operands MN-major
  {
    TcGemmDesc d = {};
    d.A = x; d.B = w.da; d.C = w.V; d.M = C; d.N = K; d.K = HW; d.lda = C; d.ldb = K; d.ldc = K;
    d.a_mn = true; d.b_mn = true; d.precision = prec;
    d.batch = B; d.sA = (long long)HW * C; d.sB = (long long)HW * K; d.sC = (long long)C * K;
    rc = tc_gemm(d, stream);
    if (rc) return rc;
  }
  nv_norm_fwd_kernel<<<B, 256, 0, stream>>>(w.V, centers, w.asum, C, w.nk, w.nt, out);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

extern "C" int scl_netvlad_bwd(const float* x, const float* assign_w, const float* centers, const float* dout, int B,
                               int HW, int C, int K, float* dx, float* dassign_w, float* dcenters, void* workspace,
                               size_t workspace_bytes, scl_stream_t stream_) {
  if (!x || !assign_w || !centers || !dout || !workspace) return SCL_ERR_BAD_ARG;
  int rc = nv_check(B, HW, C, K);
  if (rc) return rc;
  if (workspace_bytes < nv_ws_bytes(B, HW, C, K)) return SCL_ERR_WORKSPACE;
  rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NvWs w = nv_carve(workspace, workspace_bytes, B, HW, C, K);
  const long long P = (long long)B * HW;
  const int prec = tc_gemm_precision();
  if (nv_fused_ok(B, HW, C, K) && prec == 0 && workspace_bytes >= nv_ws_total(B, HW, C, K)) {
    // One kernel per image for the head (normalisations' backward -> dV, dasum, operand splits), then two kernels with one
    // pass over x each.  netvlad_fused.cu (the forward's skeleton): da = Xn dV[b], the soft-max backward, the row term of
    // the l2-normalisation backward, dW.  netvlad_dx.cu: dx = inv ([A | dS] . [dV[b] | W]^T) - rb x on operands that are
    // already fp16 hi / lo halves, through a shared-memory tile (TMA in, TMA out).
    if (dx && !aligned16(dx)) return SCL_ERR_ALIGN;
    const size_t base = nv_ws_bytes(B, HW, C, K);
    const NvBwdHead head = {w.V, dout, w.nk, w.nt, centers};
    rc = nv_fused_bwd(x, w.a, w.inv, w.dV, w.dasum, B, HW, C, w.da, w.rb, dassign_w, dx, static_cast<char*>(workspace) + base,
                      workspace_bytes - base, stream, &head);
    if (rc == SCL_OK && dcenters) {
      nv_dcenters_kernel<<<(C * 64 + 255) / 256, 256, 0, stream>>>(w.dV, w.asum, B, C, dcenters);
      SCL_LAUNCH_CHECK();
    }
    if (rc != SCL_ERR_UNSUPPORTED) return rc;        // (UNSUPPORTED is returned before anything is launched)
  }
  nv_norm_bwd_kernel<<<B, 256, 0, stream>>>(w.V, dout, w.nk, w.nt, C, w.dV);
  SCL_LAUNCH_CHECK();
  if (dcenters) {
    nv_dcenters_kernel<<<(C * 64 + 255) / 256, 256, 0, stream>>>(w.dV, w.asum, B, C, dcenters);
    SCL_LAUNCH_CHECK();
  }
  nv_dasum_kernel<<<B, 256, 0, stream>>>(w.dV, centers, C, w.dasum);
  SCL_LAUNCH_CHECK();
  // da[b] = (X[b] dV[b]) * inv[row]      M = HW, N = K, contraction over C; dV[b] [C,K] read MN-major
  {
    TcGemmDesc d = {};
    d.A = x; d.B = w.dV; d.C = w.da; d.M = HW; d.N = K; d.K = C; d.lda = C; d.ldb = K; d.ldc = K;
    d.a_mn = false; d.b_mn = true; d.rowscale = w.inv; d.rowscale_stride = HW; d.precision = prec;
    d.batch = B; d.sA = (long long)HW * C; d.sB = (long long)C * K; d.sC = (long long)HW * K;
    rc = tc_gemm(d, stream);
    if (rc) return rc;
  }
  nv_softmax_bwd_kernel<<<unsigned((P + 7) / 8), 256, 0, stream>>>(w.a, w.da, w.dasum, P, HW);
  SCL_LAUNCH_CHECK();
  if (dx) {
    if (!aligned16(dx)) return SCL_ERR_ALIGN;
    // dxh[b] = A[b] dV[b]^T (aggregation path) + dS[b] W^T (assignment path): ONE contraction over the concatenated
    // K = 64 + 64 ([A | dS] . [dV[b] | W]^T), so dx is written once instead of written, re-read and written again
    TcGemmDesc e = {};
    e.A = w.a; e.B = w.dV; e.C = dx; e.M = HW; e.N = C; e.K = K; e.lda = K; e.ldb = K; e.ldc = C;
    e.a_mn = false; e.b_mn = false; e.precision = prec;
    e.batch = B; e.sA = (long long)HW * K; e.sB = (long long)C * K; e.sC = (long long)HW * C;
    e.A2 = w.da; e.B2 = assign_w; e.K2 = K; e.sB2 = 0;
    rc = tc_gemm(e, stream);
    if (rc) return rc;
    nv_l2norm_bwd_kernel<<<unsigned((P + 7) / 8), 256, 0, stream>>>(x, w.inv, P, C, dx);
    SCL_LAUNCH_CHECK();
  }
  if (dassign_w) {
    // dW = X^T (dS * inv): per-image partials [B,C,K] (M = C, N = K, contraction over HW, both operands MN-major),
    // then a fixed-order sum over the images.  dS is scaled in place: nothing reads it afterwards.
    nv_scale_rows_kernel<<<unsigned((P + 7) / 8), 256, 0, stream>>>(w.da, P, w.inv);
    SCL_LAUNCH_CHECK();
    TcGemmDesc d = {};
    d.A = x; d.B = w.da; d.C = w.part; d.M = C; d.N = K; d.K = HW; d.lda = C; d.ldb = K; d.ldc = K;
    d.a_mn = true; d.b_mn = true; d.precision = prec;
    d.batch = B; d.sA = (long long)HW * C; d.sB = (long long)HW * K; d.sC = (long long)C * K;
    rc = tc_gemm(d, stream);
    if (rc) return rc;
    nv_sum_batch_kernel<<<(C * K + 255) / 256, 256, 0, stream>>>(w.part, B, C * K, dassign_w);
    SCL_LAUNCH_CHECK();
  }
  return SCL_OK;
}

// ---------------------------------------------------------------------------------------------
// P1: PCA whitening on the tcgen05 GEMM (tc_gemm.cu).  The centring x - m is done first, in fp32, exactly like the
// reference graph (train.py:650), so the contraction sees the same operands; 1/sqrt(var) is a column scale of the
// epilogue (forward) and a pre-scale of dy (backward: dx = (dy / sqrt(var)) V, V read MN-major, no transposed copy).
__global__ void __launch_bounds__(256) pca_center_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                         long long n4, int Din4, float* __restrict__ xc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = ldg_stream(reinterpret_cast<const float4*>(x) + i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(m) + (i % Din4));
    reinterpret_cast<float4*>(xc)[i] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
  }
}
__global__ void pca_rs_kernel(const float* __restrict__ var, int Dout, float* __restrict__ rs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Dout) rs[i] = 1.0f / sqrtf(var[i]);
}
__global__ void __launch_bounds__(256) pca_scale_dy_kernel(const float* __restrict__ dy, const float* __restrict__ var,
                                                           long long n, int Dout, float* __restrict__ dys) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dys[i] = dy[i] / sqrtf(var[i % Dout]);
}

static bool pca_tc_ok(const void* a, const void* b, const void* c, int Din, int Dout) {
  return (Din % 4 == 0) && (Dout % 4 == 0) && aligned16(a) && aligned16(b) && aligned16(c) && knob_or(KNOB_GEMM_SIMT, 0) == 0;
}

extern "C" int scl_pca_workspace_bytes(int B, int Din, int Dout, size_t* bytes) {
  if (!bytes || B < 1 || Din < 1 || Dout < 1) return SCL_ERR_BAD_ARG;
  const size_t fwd = carve_bytes(size_t(B) * Din, 4) + carve_bytes(size_t(Dout), 4);
  const size_t bwd = carve_bytes(size_t(B) * Dout, 4);
  *bytes = fwd > bwd ? fwd : bwd;
  return SCL_OK;
}

extern "C" int scl_pca_fwd(const float* x, const float* v, const float* m, const float* var, int B, int Din, int Dout,
                           float* y, void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!x || !v || !m || !var || !y) return SCL_ERR_BAD_ARG;
  if (B < 1 || Din < 1 || Dout < 1) return SCL_ERR_BAD_SHAPE;
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (pca_tc_ok(x, v, y, Din, Dout) && aligned16(m)) {
    size_t need = 0;
    scl_pca_workspace_bytes(B, Din, Dout, &need);
    if (!workspace || workspace_bytes < need) return SCL_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255u) return SCL_ERR_ALIGN;
    Carver c(workspace, workspace_bytes);
    float* xc = c.take<float>(size_t(B) * Din);
    float* rs = c.take<float>(Dout);
    const long long n4 = (long long)B * Din / 4;
    pca_center_kernel<<<unsigned(std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 16)), 256, 0, stream>>>(x, m, n4, Din / 4, xc);
    SCL_LAUNCH_CHECK();
    pca_rs_kernel<<<(Dout + 255) / 256, 256, 0, stream>>>(var, Dout, rs);
    SCL_LAUNCH_CHECK();
    TcGemmDesc d = {};
    d.A = xc; d.B = v; d.C = y; d.M = B; d.N = Dout; d.K = Din; d.lda = Din; d.ldb = Din; d.ldc = Dout;
    d.a_mn = false; d.b_mn = false; d.colscale = rs; d.precision = tc_gemm_precision();
    // few output tiles and a long K (B = 256: 2 x 32 tiles, K = 32768): two CTAs per tile fill the machine
    const int tiles = ((B + 127) / 128) * ((Dout + 127) / 128);
    if (2 * tiles <= num_sms() && Din >= 2048) {
      d.split_k = 2;
      SCL_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(B) * Dout * sizeof(float), stream));
    }
    return tc_gemm(d, stream);
  }
  // shapes the TMA path cannot address (row pitch not a multiple of 16 bytes): FP32 FFMA GEMM
  GemmArgs g = gemm_args(x, v, y, B, Dout, Din, Din, Din, Dout, 0, 1);      // (x - m) V^T, then / sqrt(var)
  g.a_sub_k = m;
  g.col_isqrt = var;
  const int ctas = ((B + 127) / 128) * ((Dout + 127) / 128);
  const int chunks = (Din + kGemmBK - 1) / kGemmBK;
  int split = (2 * num_sms() + ctas - 1) / ctas;
  if (split > chunks / 8) split = chunks / 8;
  if (split < 1) split = 1;
  g.split_k = split;
  if (split > 1) SCL_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(B) * Dout * sizeof(float), stream));
  return gemm_launch(g, stream);
}

// ---------------------------------------------------------------------------------------------
// P1 with a PREPARED projection matrix.  v (the PCA components) is a fed constant of the training loop
// (train/train.py:281-283, 647-649: the same host array every step) and of top-n.py's sweep, so it is split ONCE into
// fp16 hi / lo halves with one power-of-two scale (scl_pca_prepare -> "shadow", like the retrieval index) and every
// projection / back-projection runs on the pre-split f16 engine of tc_gemm_h3.cu: fp32-grade like the 3xTF32 path,
// 2.5-3x faster (no in-kernel split, half the operand bytes, f16 tensor rate).  The [B, Din] partner is centred in
// fp32 like the reference graph and split per row on the way.
struct PcaShadowHeader {
  float unscale;            // 2^-e of the whole matrix
  unsigned int maxbits;
  int Din, Dout, built, pad[3];
};
static size_t pca_shadow_hi_off() { return 256; }
static size_t pca_shadow_lo_off(int Din, int Dout) { return 256 + round_up(size_t(Din) * Dout * 2, 256); }

__global__ void pca_shadow_finish_kernel(PcaShadowHeader* h, int Din, int Dout) {
  h->Din = Din; h->Dout = Dout; h->built = 1;
}
__global__ void pca_colscale_kernel(const float* __restrict__ var, const PcaShadowHeader* __restrict__ h, int Dout,
                                    float* __restrict__ cs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Dout) cs[i] = h->unscale / sqrtf(var[i]);
}

extern "C" int scl_pca_shadow_bytes(int Din, int Dout, size_t* bytes) {
  if (!bytes || Din < 8 || Dout < 1 || (Din & 7)) return SCL_ERR_BAD_ARG;
  *bytes = pca_shadow_lo_off(Din, Dout) + round_up(size_t(Din) * Dout * 2, 256);
  return SCL_OK;
}

extern "C" int scl_pca_prepare(const float* v, int Din, int Dout, void* shadow, size_t shadow_bytes, scl_stream_t stream_) {
  if (!v || !shadow) return SCL_ERR_BAD_ARG;
  size_t need = 0;
  int rc = scl_pca_shadow_bytes(Din, Dout, &need);
  if (rc) return rc;
  if (shadow_bytes < need) return SCL_ERR_WORKSPACE;
  if (!aligned16(v) || (reinterpret_cast<uintptr_t>(shadow) & 255u)) return SCL_ERR_ALIGN;
  if ((rc = check_device())) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  char* sb = static_cast<char*>(shadow);
  PcaShadowHeader* h = reinterpret_cast<PcaShadowHeader*>(sb);
  rc = h3_split_all(v, (long long)Din * Dout, &h->maxbits, reinterpret_cast<__half*>(sb + pca_shadow_hi_off()),
                    reinterpret_cast<__half*>(sb + pca_shadow_lo_off(Din, Dout)), &h->unscale, stream);
  if (rc) return rc;
  pca_shadow_finish_kernel<<<1, 1, 0, stream>>>(h, Din, Dout);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

extern "C" int scl_pca_prepared_workspace_bytes(int B, int Din, int Dout, size_t* bytes) {
  if (!bytes || B < 1 || Din < 1 || Dout < 1) return SCL_ERR_BAD_ARG;
  const int W = Din > Dout ? Din : Dout;
  *bytes = 2 * carve_bytes(size_t(B) * W, 2) + carve_bytes(size_t(B), 4) + carve_bytes(size_t(Dout), 4);
  return SCL_OK;
}

static bool pca_prepared_ok(int B, int Din, int Dout) {
  return B >= 1 && Din >= 8 && (Din & 7) == 0 && (Dout & 7) == 0 && Din <= 32768 && Dout <= 32768;
}

extern "C" int scl_pca_fwd_prepared(const float* x, const void* shadow, const float* m, const float* var, int B, int Din,
                                    int Dout, float* y, void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!x || !shadow || !m || !var || !y || !workspace) return SCL_ERR_BAD_ARG;
  if (!pca_prepared_ok(B, Din, Dout)) return SCL_ERR_UNSUPPORTED;
  size_t need = 0;
  scl_pca_prepared_workspace_bytes(B, Din, Dout, &need);
  if (workspace_bytes < need) return SCL_ERR_WORKSPACE;
  if (!aligned16(x) || !aligned16(y) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return SCL_ERR_ALIGN;
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const char* sb = static_cast<const char*>(shadow);
  const PcaShadowHeader* h = reinterpret_cast<const PcaShadowHeader*>(sb);
  const __half* vh = reinterpret_cast<const __half*>(sb + pca_shadow_hi_off());
  const __half* vl = reinterpret_cast<const __half*>(sb + pca_shadow_lo_off(Din, Dout));
  const int W = Din > Dout ? Din : Dout;
  Carver c(workspace, workspace_bytes);
  __half* xh = c.take<__half>(size_t(B) * W);
  __half* xl = c.take<__half>(size_t(B) * W);
  float* un = c.take<float>(B);
  float* cs = c.take<float>(Dout);
  if ((rc = h3_split_rows(x, m, nullptr, B, Din, xh, xl, un, nullptr, stream))) return rc;      // (x - m), train.py:650
  pca_colscale_kernel<<<(Dout + 255) / 256, 256, 0, stream>>>(var, h, Dout, cs);              // / sqrt(var), :651
  SCL_LAUNCH_CHECK();
  const int tiles = ((B + 127) / 128) * ((Dout + 127) / 128);
  const int split = (2 * tiles <= num_sms() && Din >= 2048) ? 2 : 1;
  if (split == 2) SCL_CUDA_TRY(cudaMemsetAsync(y, 0, size_t(B) * Dout * sizeof(float), stream));
  return tc_gemm_h3(xh, xl, vh, vl, y, B, Dout, Din, Din, Din, Dout, false, un, cs, split, stream);
}

extern "C" int scl_pca_bwd_prepared(const float* dy, const void* shadow, const float* var, int B, int Din, int Dout,
                                    float* dx, void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!dy || !shadow || !var || !dx || !workspace) return SCL_ERR_BAD_ARG;
  if (!pca_prepared_ok(B, Din, Dout)) return SCL_ERR_UNSUPPORTED;
  size_t need = 0;
  scl_pca_prepared_workspace_bytes(B, Din, Dout, &need);
  if (workspace_bytes < need) return SCL_ERR_WORKSPACE;
  if (!aligned16(dy) || !aligned16(dx) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return SCL_ERR_ALIGN;
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const char* sb = static_cast<const char*>(shadow);
  const PcaShadowHeader* h = reinterpret_cast<const PcaShadowHeader*>(sb);
  const __half* vh = reinterpret_cast<const __half*>(sb + pca_shadow_hi_off());
  const __half* vl = reinterpret_cast<const __half*>(sb + pca_shadow_lo_off(Din, Dout));
  const int W = Din > Dout ? Din : Dout;
  Carver c(workspace, workspace_bytes);
  __half* dh = c.take<__half>(size_t(B) * W);
  __half* dl = c.take<__half>(size_t(B) * W);
  float* un = c.take<float>(B);
  // dx = (dy / sqrt(var)) V: rows of dy / sqrt(var) split per row (the matrix scale folded into the row scale), V read
  // MN-major (contraction over its rows) from the same shadow
  if ((rc = h3_split_rows(dy, nullptr, var, B, Dout, dh, dl, un, &h->unscale, stream))) return rc;
  return tc_gemm_h3(dh, dl, vh, vl, dx, B, Din, Dout, Dout, Din, Din, true, un, nullptr, 1, stream);
}

// ---------------------------------------------------------------------------------------------
// P0: the data-sized steps of PCA(whiten=True).fit (evaluation/top-n.py:74-75): column means
// and the centred copy.  The contractions of the fit (Gram X_c X_c^T or covariance X_c^T X_c, and the back-projection
// X_c^T U) run on scl_gemm_tf32; the small symmetric eigenproblem stays with the caller.
// Column sums are accumulated in float64 per row slice (coalesced: a warp reads 32 consecutive columns of a row) and
// the slices are added in a fixed order: deterministic.
constexpr int kPcaSlices = 16;
__global__ void __launch_bounds__(256) pca_colsum_kernel(const float* __restrict__ x, int n, int D, double* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  const int rows = (n + kPcaSlices - 1) / kPcaSlices;
  const int r0 = blockIdx.y * rows, r1 = min(n, r0 + rows);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    const float v0 = ldg_stream(x + size_t(r) * D + c), v1 = ldg_stream(x + size_t(r + 1) * D + c);
    const float v2 = ldg_stream(x + size_t(r + 2) * D + c), v3 = ldg_stream(x + size_t(r + 3) * D + c);
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
  }
  for (; r < r1; ++r) a0 += ldg_stream(x + size_t(r) * D + c);
  part[size_t(blockIdx.y) * D + c] = (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(256) pca_mean_kernel(const double* __restrict__ part, int n, int D, float* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  double a = 0.0;
#pragma unroll
  for (int s = 0; s < kPcaSlices; ++s) a += part[size_t(s) * D + c];
  mean[c] = float(a / double(n));
}
__global__ void __launch_bounds__(256) pca_center_any_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                             long long total, int D, float* __restrict__ xc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    xc[i] = ldg_stream(x + i) - __ldg(m + (i % D));
}

extern "C" int scl_pca_center_workspace_bytes(int n, int D, size_t* bytes) {
  if (!bytes || n < 1 || D < 1) return SCL_ERR_BAD_ARG;
  *bytes = carve_bytes(size_t(kPcaSlices) * D, sizeof(double));
  return SCL_OK;
}

extern "C" int scl_pca_center(const float* x, int n, int D, float* mean, float* xc, void* workspace, size_t workspace_bytes,
                              scl_stream_t stream_) {
  if (!x || !mean) return SCL_ERR_BAD_ARG;
  if (n < 1 || D < 1) return SCL_ERR_BAD_SHAPE;
  int rc = check_device();
  if (rc) return rc;
  size_t need = 0;
  scl_pca_center_workspace_bytes(n, D, &need);
  if (!workspace || workspace_bytes < need) return SCL_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return SCL_ERR_ALIGN;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver c(workspace, workspace_bytes);
  double* part = c.take<double>(size_t(kPcaSlices) * D);
  pca_colsum_kernel<<<dim3((D + 255) / 256, kPcaSlices), 256, 0, stream>>>(x, n, D, part);
  SCL_LAUNCH_CHECK();
  pca_mean_kernel<<<(D + 255) / 256, 256, 0, stream>>>(part, n, D, mean);
  SCL_LAUNCH_CHECK();
  if (xc) {
    const long long total = (long long)n * D;
    if (D % 4 == 0 && aligned16(x) && aligned16(xc) && aligned16(mean)) {
      const long long n4 = total / 4;
      pca_center_kernel<<<unsigned(std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 16)), 256, 0, stream>>>(x, mean, n4, D / 4, xc);
    } else {
      pca_center_any_kernel<<<unsigned(std::min<long long>((total + 255) / 256, (long long)num_sms() * 16)), 256, 0, stream>>>(x, mean, total, D, xc);
    }
    SCL_LAUNCH_CHECK();
  }
  return SCL_OK;
}

extern "C" int scl_pca_bwd(const float* dy, const float* v, const float* var, int B, int Din, int Dout, float* dx,
                           void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!dy || !v || !var || !dx) return SCL_ERR_BAD_ARG;
  if (B < 1 || Din < 1 || Dout < 1) return SCL_ERR_BAD_SHAPE;
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (pca_tc_ok(dy, v, dx, Din, Dout)) {
    size_t need = 0;
    scl_pca_workspace_bytes(B, Din, Dout, &need);
    if (!workspace || workspace_bytes < need) return SCL_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255u) return SCL_ERR_ALIGN;
    float* dys = static_cast<float*>(workspace);
    const long long n = (long long)B * Dout;
    pca_scale_dy_kernel<<<unsigned(std::min<long long>((n + 255) / 256, (long long)num_sms() * 16)), 256, 0, stream>>>(dy, var, n, Dout, dys);
    SCL_LAUNCH_CHECK();
    TcGemmDesc d = {};
    d.A = dys; d.B = v; d.C = dx; d.M = B; d.N = Din; d.K = Dout; d.lda = Dout; d.ldb = Din; d.ldc = Din;
    d.a_mn = false; d.b_mn = true; d.colscale = nullptr; d.precision = tc_gemm_precision();
    return tc_gemm(d, stream);
  }
  GemmArgs g = gemm_args(dy, v, dx, B, Din, Dout, Dout, Din, Din, 0, 0);    // (dy / sqrt(var)) V
  g.a_isqrt_k = var;
  return gemm_launch(g, stream);
}
