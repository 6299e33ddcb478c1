// flat_losses.cu -- W1 / W2 flat mode: one Gram matrix over the whole batch.
//
// Replaces wms_loss with 2-D inputs (/root/reference/model/losses.py:5-60) and ms_loss (:76-122, labels from
// train/train.py:821-826) plus their autodiff backward, for B descriptors of dimension D:
//   1. G = E E^T                                   (sgemm.cuh, FP32)                      losses.py:25 / :94
//   2. one CTA per anchor row: similarities, masks, mining thresholds, log-sum-exp weights (ms_row.cuh)
//      -> Gw = dL/dS, row losses, optional kept-pair bitmaps                              losses.py:26-58
//   3. M = diag(1/|e|) (Gw + Gw^T - diag(c)) diag(1/|e|)   (l2_normalize Jacobian folded in)
//   4. dE = M E                                    (sgemm.cuh)
#include "ms_row.cuh"
#include "sgemm.cuh"
#include "tc_gemm.cuh"

namespace scl {

struct FlatWs {
  float* G;        // [B,B] raw Gram
  float* Gw;       // [B,B] dL/dS
  float* Mm;       // [B,B]
  float* invn;     // [B]
  float* nflag;    // [B]
  float* rowloss;  // [B]
  float* cvec;     // [B]
};

static size_t flat_ws_bytes(int B) {
  return 3 * carve_bytes(size_t(B) * B, sizeof(float)) + 4 * carve_bytes(B, sizeof(float));
}
static FlatWs flat_carve(void* ws, size_t bytes, int B) {
  Carver c(ws, bytes);
  FlatWs w;
  w.G = c.take<float>(size_t(B) * B);
  w.Gw = c.take<float>(size_t(B) * B);
  w.Mm = c.take<float>(size_t(B) * B);
  w.invn = c.take<float>(B);
  w.nflag = c.take<float>(B);
  w.rowloss = c.take<float>(B);
  w.cvec = c.take<float>(B);
  return w;
}

__global__ void flat_norm_kernel(const float* __restrict__ G, int B, float* __restrict__ invn,
                                 float* __restrict__ nflag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    float n2 = G[size_t(i) * B + i];
    invn[i] = rsqrtf(fmaxf(n2, 1e-12f));          // tf.nn.l2_normalize (losses.py:7 / :84)
    nflag[i] = n2 >= 1e-12f ? 1.0f : 0.0f;
  }
}

__device__ __forceinline__ float block_reduce(float v, float* sh, int op) {   // op 0 sum, 1 max, 2 min
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = op == 0 ? warp_sum(v) : (op == 1 ? warp_max(v) : warp_min(v));
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = op == 0 ? 0.0f : (op == 1 ? -INFINITY : INFINITY);
  for (int w = 0; w < nw; ++w) r = op == 0 ? r + sh[w] : (op == 1 ? fmaxf(r, sh[w]) : fminf(r, sh[w]));
  return r;
}

// One CTA per anchor row i.  dist == nullptr selects ms_loss (hard masks from labels).
__global__ void __launch_bounds__(256) flat_row_kernel(const float* __restrict__ G, const float* __restrict__ dist,
                                                       const int32_t* __restrict__ labels, int B, scl_ms_params p,
                                                       const float* __restrict__ invn, float* __restrict__ Gw,
                                                       float* __restrict__ rowloss, uint8_t* __restrict__ kept) {
  extern __shared__ float sh[];
  float* s_s = sh;            // clamped similarity
  float* s_wp = sh + B;
  float* s_wn = sh + 2 * B;
  float* s_red = sh + 3 * B;  // 32 floats
  const int i = blockIdx.x;
  const float inv_i = invn[i];
  const int li = labels ? labels[i] : 0;
  float mx_neg = -INFINITY, mx_pos = -INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float raw = G[size_t(i) * B + j] * inv_i * invn[j];
    float s = fmaxf(raw, 0.0f);                                   // losses.py:26 / :95
    float wp, wn;
    if (dist) {
      wms_masks(dist[size_t(i) * B + j], p.d_alpha, p.d_beta, p.wfunction, wp, wn);
    } else {
      bool same = labels[j] == li;                                // losses.py:89-93
      wp = same ? 1.0f : 0.0f;
      wn = same ? 0.0f : 1.0f;
    }
    if (i == j) wp -= 1.0f;                                       // losses.py:22 / :92
    s_s[j] = s;
    s_wp[j] = wp;
    s_wn[j] = wn;
    mx_neg = fmaxf(mx_neg, s * wn);
    mx_pos = fmaxf(mx_pos, s * wp);
  }
  MsRowStats st;
  st.maxv = block_reduce(mx_neg, s_red, 1);
  st.tmp = block_reduce(mx_pos, s_red, 1);
  float mn = INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) mn = fminf(mn, (s_s[j] - st.tmp) * s_wp[j]);
  st.minv = block_reduce(mn, s_red, 2) + st.tmp;
  float A = 0.0f, Bn = 0.0f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    bool kp, kn;
    float ep, en;
    ms_elem(s_s[j], s_wp[j], s_wn[j], st, p, kp, kn, ep, en);
    A += ep;
    Bn += en;
  }
  A = block_reduce(A, s_red, 0);
  Bn = block_reduce(Bn, s_red, 0);
  const float invB = 1.0f / float(B);
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    bool kp, kn;
    float ep, en;
    ms_elem(s_s[j], s_wp[j], s_wn[j], st, p, kp, kn, ep, en);
    float g = ms_elem_grad(s_wp[j], s_wn[j], kp, kn, ep, en, A, Bn, p) * invB;
    float raw = G[size_t(i) * B + j] * inv_i * invn[j];
    if (!(raw >= 0.0f)) g = 0.0f;                                 // tf.maximum gradient gate
    Gw[size_t(i) * B + j] = g;
    if (kept) {
      kept[size_t(i) * B + j] = kp ? 1 : 0;
      kept[size_t(B) * B + size_t(i) * B + j] = kn ? 1 : 0;
    }
  }
  if (threadIdx.x == 0) rowloss[i] = ms_row_loss(A, Bn, p) * invB;
}

// c_i = nflag_i * sum_j (Gw_ij + Gw_ji) * raw_ij ;  M_ij = invn_i (Gw_ij + Gw_ji - c_i [i==j]) invn_j
__global__ void __launch_bounds__(256) flat_m_kernel(const float* __restrict__ G, const float* __restrict__ Gw, int B,
                                                     const float* __restrict__ invn, const float* __restrict__ nflag,
                                                     float* __restrict__ Mm) {
  __shared__ float s_red[32];
  const int i = blockIdx.x;
  const float inv_i = invn[i];
  float part = 0.0f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float w = Gw[size_t(i) * B + j] + Gw[size_t(j) * B + i];
    part += w * (G[size_t(i) * B + j] * inv_i * invn[j]);
  }
  const float c = block_reduce(part, s_red, 0) * nflag[i];
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float w = Gw[size_t(i) * B + j] + Gw[size_t(j) * B + i];
    if (i == j) w -= c;
    Mm[size_t(i) * B + j] = inv_i * w * invn[j];
  }
}

__global__ void __launch_bounds__(256) flat_loss_kernel(const float* __restrict__ rowloss, int B,
                                                        float* __restrict__ loss) {
  __shared__ float s_red[32];
  float v = 0.0f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) v += rowloss[j];
  v = block_reduce(v, s_red, 0);
  if (threadIdx.x == 0) loss[0] = v;
}

static int flat_run(const float* emb, const float* dist, const int32_t* labels, int B, int D, const scl_ms_params* p,
                    float* loss, float* demb, uint8_t* kept, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
  if (!emb || !p || !loss || !workspace || (!dist && !labels)) return SCL_ERR_BAD_ARG;
  if (B < 2 || D < 1 || B > 16384) return SCL_ERR_BAD_SHAPE;
  if (!aligned16(workspace)) return SCL_ERR_ALIGN;
  int rc = check_device();
  if (rc) return rc;
  if (workspace_bytes < flat_ws_bytes(B)) return SCL_ERR_WORKSPACE;
  FlatWs w = flat_carve(workspace, workspace_bytes, B);

  // G = E E^T and dE = M E run on the tcgen05 GEMM (fp32-grade 3xTF32 unless scl_set_gemm_precision(1)); shapes
  // whose row pitch TMA cannot address (B or D not a multiple of 4) use the FP32 FFMA GEMM
  const bool tc = (B % 4 == 0) && (D % 4 == 0) && aligned16(emb) && (!demb || aligned16(demb)) && knob_or(KNOB_GEMM_SIMT, 0) == 0;
  if (tc) {
    TcGemmDesc d = {};
    d.A = emb; d.B = emb; d.C = w.G; d.M = B; d.N = B; d.K = D; d.lda = D; d.ldb = D; d.ldc = B;
    d.a_mn = false; d.b_mn = false; d.colscale = nullptr; d.precision = tc_gemm_precision();
    // few output tiles and a long K (B = 1024: 64 tiles, K = 4096): two CTAs per tile fill the machine; the two halves
    // meet in one atomicAdd per element (two addends commute: deterministic)
    const int tiles = ((B + 127) / 128) * ((B + 127) / 128);
    if (2 * tiles <= num_sms() && D >= 2048) {
      d.split_k = 2;
      SCL_CUDA_TRY(cudaMemsetAsync(w.G, 0, size_t(B) * B * sizeof(float), stream));
    }
    rc = tc_gemm(d, stream);
  } else {
    GemmArgs g = gemm_args(emb, emb, w.G, B, B, D, D, D, B, 0, 1);       // G = E E^T
    rc = gemm_launch(g, stream);
  }
  if (rc) return rc;
  flat_norm_kernel<<<(B + 255) / 256, 256, 0, stream>>>(w.G, B, w.invn, w.nflag);
  SCL_LAUNCH_CHECK();
  const size_t shm = (3 * size_t(B) + 32) * sizeof(float);
  if (shm > 48 * 1024) {
    SCL_CUDA_TRY(cudaFuncSetAttribute(flat_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shm)));
  }
  flat_row_kernel<<<B, 256, shm, stream>>>(w.G, dist, labels, B, *p, w.invn, w.Gw, w.rowloss, kept);
  SCL_LAUNCH_CHECK();
  flat_loss_kernel<<<1, 256, 0, stream>>>(w.rowloss, B, loss);
  SCL_LAUNCH_CHECK();
  if (demb) {
    flat_m_kernel<<<B, 256, 0, stream>>>(w.G, w.Gw, B, w.invn, w.nflag, w.Mm);
    SCL_LAUNCH_CHECK();
    if (tc) {
      TcGemmDesc d = {};
      d.A = w.Mm; d.B = emb; d.C = demb; d.M = B; d.N = D; d.K = B; d.lda = B; d.ldb = D; d.ldc = D;
      d.a_mn = false; d.b_mn = true; d.colscale = nullptr; d.precision = tc_gemm_precision();   // E read MN-major
      rc = tc_gemm(d, stream);
    } else {
      GemmArgs h = gemm_args(w.Mm, emb, demb, B, D, B, B, D, D, 0, 0);   // dE = M E
      rc = gemm_launch(h, stream);
    }
    if (rc) return rc;
  }
  return SCL_OK;
}

}  // namespace scl

extern "C" int scl_ms_flat_workspace_bytes(int B, int D, size_t* bytes) {
  (void)D;
  if (!bytes || B < 2) return SCL_ERR_BAD_ARG;
  *bytes = scl::flat_ws_bytes(B);
  return SCL_OK;
}

extern "C" int scl_wms_flat_fwd_bwd(const float* emb, const float* dist, int B, int D, const scl_ms_params* p,
                                    float* loss, float* demb, uint8_t* kept, void* workspace, size_t workspace_bytes,
                                    scl_stream_t stream) {
  if (!dist) return SCL_ERR_BAD_ARG;
  return scl::flat_run(emb, dist, nullptr, B, D, p, loss, demb, kept, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int scl_ms_flat_fwd_bwd(const float* emb, const int32_t* labels, int B, int D, const scl_ms_params* p,
                                   float* loss, float* demb, uint8_t* kept, void* workspace, size_t workspace_bytes,
                                   scl_stream_t stream) {
  if (!labels || !p) return SCL_ERR_BAD_ARG;
  scl_ms_params q = *p;
  q.sumfunction = SCL_SUM_MS;     // ms_loss has only the 'ms' sum (losses.py:111-120)
  return scl::flat_run(emb, nullptr, labels, B, D, &q, loss, demb, kept, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}
