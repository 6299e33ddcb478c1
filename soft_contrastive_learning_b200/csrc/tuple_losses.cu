// tuple_losses.cu -- L1-L6 (triplet family, Huber-distance triplet, log-ratio) and D1 (pairwise squared distances).
//
// Replaces, forward + analytic backward:
//   pointnetvlad_cls.{triplet,lazy_triplet,quadruplet,lazy_quadruplet}_loss  (train/train.py:700-712)
//   evil_triplet_loss / evil_quadruplet_loss                                  (model/losses.py:63-73,197-214)
//   distance_triplet_loss with distance_loss / huber_distance_loss            (model/losses.py:225-264,678-690)
//   distance_quadruplet_loss, same two distance terms x {triplet, lazy}       (model/losses.py:267-307,664-675)
//   logratio_loss                                                             (model/losses.py:125-135)
//   _pairwise_squared_distances                                               (model/losses.py:656-661)
//
// All of these only need, per tuple, the squared distances of every row to the anchor (row 0) and to the
// `other` negative (last row) -- computed as direct differences like the reference's squared_difference, not via
// the Gram identity -- followed by O(S) scalar logic.  The gradient of every row is a linear combination of the
// tuple's rows, so the backward is the same M * E product over the shared-memory-resident slice as in
// wms_tuple.cu (tuple_common.cuh).  One cluster per tuple, one HBM read and one HBM write per descriptor.
#include <atomic>
#include <cstdlib>

#include "tuple_common.cuh"

namespace scl {

enum { kModeTriplet = 0, kModeLogratio = 1 };

struct AnchorArgs {
  int mode;
  int P, N, has_other;
  scl_tuple_params tp;
  const float* sq_d_dists;   // [T,P]    (distance term)
  const float* sq_pos;       // [T,P]    (logratio)
  const float* sq_neg;       // [T,N]
  int strict_reference;
};

template <int SG>
struct AnchorSmem {
  static __host__ __device__ size_t floats(int chunk_cols) {
    size_t n = 0;
    n += size_t(SG) * (chunk_cols + 4);       // Es
    n += al4(size_t(kMaxCluster) * 2 * SG);   // slots: [C][2][SG] partial distances
    n += al4(size_t(4) * SG);                 // da, dob, gamma, omega
    n += al4(size_t(SG) * SG);                // Mraw
    n += size_t(TupDims<SG>::MT);             // Mt
    n += 8;
    return n;
  }
};

// Scalar logic of one tuple, executed by warp 0.  da[j] = |e_j - e_0|^2, dob[j] = |e_j - e_last|^2.
// Writes gamma[j] = d L_t / d da[j], omega[j] = d L_t / d dob[j]; returns the tuple's loss.
__device__ float anchor_logic(const AnchorArgs& a, int t, int S, const float* da, const float* dob, float* gamma,
                              float* omega, int lane) {
  const int P = a.P, N = a.N;
  for (int j = lane; j < S; j += 32) { gamma[j] = 0.0f; omega[j] = 0.0f; }
  __syncwarp();
  float loss = 0.0f;
  if (a.mode == kModeTriplet) {
    const int kind = a.tp.kind;
    const bool evil = kind == SCL_EVIL_TRIPLET || kind == SCL_EVIL_QUADRUPLET;
    const bool lazy = kind == SCL_LAZY_TRIPLET || kind == SCL_LAZY_QUADRUPLET || kind == SCL_DISTANCE_LAZY_QUADRUPLET;
    const bool quad = kind == SCL_QUADRUPLET || kind == SCL_LAZY_QUADRUPLET || kind == SCL_EVIL_QUADRUPLET;
    const bool dquad = kind == SCL_DISTANCE_QUADRUPLET || kind == SCL_DISTANCE_LAZY_QUADRUPLET;
    // best (min) / worst (max) positive distance, ties share the gradient evenly (tf.reduce_min/max)
    float dp = (lane < P) ? da[1 + lane] : (evil ? -INFINITY : INFINITY);
    float ref = evil ? warp_max(dp) : warp_min(dp);
    unsigned tie = __ballot_sync(0xffffffffu, lane < P && dp == ref);
    float tie_w = 1.0f / float(__popc(tie));
    float dref = 0.0f;   // d L_t / d ref
    // first hinge: anchor vs negatives
    {
      float x = (lane < N) ? a.tp.m1 + ref - da[1 + P + lane] : -INFINITY;
      float h = (lane < N) ? fmaxf(x, 0.0f) : (lazy ? -INFINITY : 0.0f);
      float coef = 0.0f;
      if (lazy) {
        float hm = warp_max(h);
        unsigned tm = __ballot_sync(0xffffffffu, lane < N && h == hm);
        coef = (lane < N && h == hm && x >= 0.0f) ? 1.0f / float(__popc(tm)) : 0.0f;
        loss += hm;
      } else {
        coef = (lane < N && x >= 0.0f) ? 1.0f : 0.0f;
        loss += warp_sum(h);
      }
      if (lane < N) gamma[1 + P + lane] -= coef;
      dref += warp_sum(coef);
    }
    if (quad) {  // second hinge: other negative vs negatives
      float x = (lane < N) ? a.tp.m2 + ref - dob[1 + P + lane] : -INFINITY;
      float h = (lane < N) ? fmaxf(x, 0.0f) : (lazy ? -INFINITY : 0.0f);
      float coef = 0.0f;
      if (lazy) {
        float hm = warp_max(h);
        unsigned tm = __ballot_sync(0xffffffffu, lane < N && h == hm);
        coef = (lane < N && h == hm && x >= 0.0f) ? 1.0f / float(__popc(tm)) : 0.0f;
        loss += hm;
      } else {
        coef = (lane < N && x >= 0.0f) ? 1.0f : 0.0f;
        loss += warp_sum(h);
      }
      if (lane < N) omega[1 + P + lane] -= coef;
      dref += warp_sum(coef);
    }
    if (lane < P && dp == ref) gamma[1 + lane] += dref * tie_w;
    // distance term (losses.py:225-236): mean over positives of (f/fmax - d/dmax)^2 or Huber(delta=1)
    if (a.tp.dist_term != SCL_DIST_NONE) {
      float term = 0.0f, g = 0.0f;
      if (lane < P) {
        float sd = a.sq_d_dists[size_t(t) * P + lane] / a.tp.d_max_squared;
        float sf = da[1 + lane] / a.tp.f_max_squared;
        float e = sf - sd;
        if (a.tp.dist_term == SCL_DIST_HUBER) {
          float ae = fabsf(e), q = fminf(ae, 1.0f);
          term = 0.5f * q * q + (ae - q);
          g = (ae <= 1.0f) ? e : (e > 0.0f ? 1.0f : -1.0f);
        } else {
          term = e * e;
          g = 2.0f * e;
        }
        g = a.tp.lam * g / (a.tp.f_max_squared * float(P));
        gamma[1 + lane] += g;
      }
      loss += a.tp.lam * warp_sum(term) / float(P);
    }
    // distance_quadruplet_loss (losses.py:267-307): the second hinge uses the smallest element of the distance term as
    // "best positive" (losses.py:664-675), scales |neg - other|^2 by f_max and always takes the max over negatives
    if (dquad) {
      float el = INFINITY, gel = 0.0f;
      if (lane < P) {
        const float sd = a.sq_d_dists[size_t(t) * P + lane] / a.tp.d_max_squared;
        const float e = da[1 + lane] / a.tp.f_max_squared - sd;
        if (a.tp.dist_term == SCL_DIST_HUBER) {
          const float ae = fabsf(e), q = fminf(ae, 1.0f);
          el = 0.5f * q * q + (ae - q);
          gel = (ae <= 1.0f) ? e : (e > 0.0f ? 1.0f : -1.0f);
        } else {
          el = e * e;
          gel = 2.0f * e;
        }
      }
      const float best = warp_min(el);
      const unsigned btie = __ballot_sync(0xffffffffu, lane < P && el == best);
      const float x = (lane < N) ? a.tp.m2 + best - dob[1 + P + lane] / a.tp.f_max_squared : -INFINITY;
      const float hh = (lane < N) ? fmaxf(x, 0.0f) : -INFINITY;
      const float hm = warp_max(hh);
      const unsigned tm = __ballot_sync(0xffffffffu, lane < N && hh == hm);
      const float coef = (lane < N && hh == hm && x >= 0.0f) ? 1.0f / float(__popc(tm)) : 0.0f;
      loss += hm;
      if (lane < N) omega[1 + P + lane] -= coef / a.tp.f_max_squared;
      const float dbest = warp_sum(coef);
      if (lane < P && el == best) gamma[1 + lane] += dbest * gel / (a.tp.f_max_squared * float(__popc(btie)));
    }
  } else {
    // logratio_loss, losses.py:125-135 (T=1 formula).  fr[n][p] = log(dpos_p / dneg_n).
    // strict: dr[n] = log(sq_pos[n] / sq_neg[n]) (element-wise, P==N), broadcast along p.
    // loss = mean_{n,p} (fr[n][p] - dr)^2
    const float inv = 1.0f / float(P * N);
    float acc = 0.0f;
    for (int n = 0; n < N; ++n) {
      float dn = da[1 + P + n];
      float gneg = 0.0f;
      float drs = a.strict_reference ? logf(a.sq_pos[size_t(t) * P + n] / a.sq_neg[size_t(t) * N + n]) : 0.0f;
      if (lane < P) {
        float dpv = da[1 + lane];
        float dr = a.strict_reference ? drs : logf(a.sq_pos[size_t(t) * P + lane] / a.sq_neg[size_t(t) * N + n]);
        float diff = logf(dpv / dn) - dr;
        acc += diff * diff;
        float g = 2.0f * diff * inv;
        gamma[1 + lane] += g / dpv;     // d log(dp/dn) / d dp
        gneg = -g / dn;
      }
      gneg = warp_sum(gneg);
      if (lane == 0) gamma[1 + P + n] += gneg;
    }
    loss = warp_sum(acc) * inv;
  }
  __syncwarp();
  return loss;
}

template <int SG>
__global__ void __launch_bounds__(kTupThreads, 2) anchor_tuple_kernel(
    const float* __restrict__ emb, int T, int S, int D, int Ds, int Dc, AnchorArgs a, float* __restrict__ demb,
    float* __restrict__ loss_out, unsigned int* __restrict__ ws) {
  using Dm = TupDims<SG>;
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int crank = cluster.block_rank();
  const int t = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pitch = Dc + 4, ncols4 = Dc >> 2, nchunks = Ds / Dc;

  float* Es = smem;
  float* slots = Es + size_t(SG) * pitch;
  float* da = slots + al4(kMaxCluster * 2 * SG);
  float* dob = da + SG;
  float* gamma = dob + SG;
  float* omega = gamma + SG;
  float* Mraw = da + al4(4 * SG);
  float* Mt = Mraw + al4(SG * SG);      // 16-byte aligned: read as float4

  const float* E_t = emb + (size_t(t) * S) * D + size_t(crank) * Ds;
  if (C > 1) cluster_arrive();
  const int last = S - 1;

  // ---- distances of every row to the anchor and to the last row, over this CTA's slice ----
  float pa[(SG + kTupWarps - 1) / kTupWarps], po[(SG + kTupWarps - 1) / kTupWarps];
#pragma unroll
  for (int k = 0; k < (SG + kTupWarps - 1) / kTupWarps; ++k) { pa[k] = 0.0f; po[k] = 0.0f; }
  for (int ch = 0; ch < nchunks; ++ch) {
    tup_load_chunk(Es, E_t + size_t(ch) * Dc, S, D, Dc);
#pragma unroll
    for (int k = 0; k < (SG + kTupWarps - 1) / kTupWarps; ++k) {
      const int j = warp + k * kTupWarps;
      if (j < S) {
        float sa = 0.0f, so = 0.0f;
        for (int c4 = lane; c4 < ncols4; c4 += 32) {
          const float4 e = *reinterpret_cast<const float4*>(Es + j * pitch + 4 * c4);
          const float4 q = *reinterpret_cast<const float4*>(Es + 4 * c4);
          const float4 o = *reinterpret_cast<const float4*>(Es + last * pitch + 4 * c4);
          float d;
          d = e.x - q.x; sa = fmaf(d, d, sa);
          d = e.y - q.y; sa = fmaf(d, d, sa);
          d = e.z - q.z; sa = fmaf(d, d, sa);
          d = e.w - q.w; sa = fmaf(d, d, sa);
          d = e.x - o.x; so = fmaf(d, d, so);
          d = e.y - o.y; so = fmaf(d, d, so);
          d = e.z - o.z; so = fmaf(d, d, so);
          d = e.w - o.w; so = fmaf(d, d, so);
        }
        pa[k] += sa;
        po[k] += so;
      }
    }
    if (nchunks > 1) __syncthreads();
  }
  if (C > 1) cluster_wait();
#pragma unroll
  for (int k = 0; k < (SG + kTupWarps - 1) / kTupWarps; ++k) {
    const int j = warp + k * kTupWarps;
    float sa = warp_sum(pa[k]), so = warp_sum(po[k]);
    if (j < S && lane < C) {
      float* remote = C > 1 ? cluster.map_shared_rank(slots, lane) : slots;
      remote[crank * 2 * SG + j] = sa;
      remote[crank * 2 * SG + SG + j] = so;
    }
  }
  if (C > 1) cluster.sync(); else __syncthreads();
  if (tid < 2 * SG) {
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += slots[c * 2 * SG + tid];
    da[tid] = s;   // da and dob are contiguous
  }
  __syncthreads();

  // ---- scalar logic (warp 0), then the coefficient matrix ----
  if (warp == 0) {
    float v = anchor_logic(a, t, S, da, dob, gamma, omega, lane);
    if (crank == 0) tup_finish_loss(ws, t, T, v, loss_out, lane);
  }
  for (int k = tid; k < SG * SG; k += kTupThreads) Mraw[k] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    // d/d e_j += 2 g_j (e_j - e_0) + 2 w_j (e_j - e_L);  d/d e_0 -= 2 g_j (e_j - e_0);  d/d e_L -= 2 w_j (e_j - e_L)
    // (assembled serially by one thread: S is tiny and the summation order stays deterministic)
    for (int j = 0; j < S; ++j) {
      const float g = 2.0f * gamma[j], w = 2.0f * omega[j];
      Mraw[j * SG + j] += g + w;
      Mraw[j * SG + 0] -= g;
      Mraw[j * SG + last] -= w;
      Mraw[0 * SG + j] -= g;
      Mraw[0 * SG + 0] += g;
      Mraw[last * SG + j] -= w;
      Mraw[last * SG + last] += w;
    }
  }
  __syncthreads();
  tup_store_Mt<SG>(Mt, Mraw, SG, S);

  if (demb != nullptr) {
    float* dE_t = demb + (size_t(t) * S) * D + size_t(crank) * Ds;
    for (int ch = 0; ch < nchunks; ++ch) {
      if (nchunks > 1) tup_load_chunk(Es, E_t + size_t(ch) * Dc, S, D, Dc);
      tup_bwd_chunk<SG>(Es, Mt, dE_t + size_t(ch) * Dc, S, D, Dc, 1.0f / float(T));
      if (nchunks > 1) __syncthreads();
    }
  }
}

static size_t anchor_smem_bytes(int sg, int dc) {
  switch (sg) {
    case 25: return AnchorSmem<25>::floats(dc) * sizeof(float);
    case 30: return AnchorSmem<30>::floats(dc) * sizeof(float);
    default: return AnchorSmem<35>::floats(dc) * sizeof(float);
  }
}

static int anchor_plan(int S, int D, TupPlan* pl) {
  if (S < 2 || S > 35 || D < 4 || (D & 3)) return SCL_ERR_BAD_SHAPE;
  pl->sg = S <= 25 ? 25 : (S <= 30 ? 30 : 35);
  int c = 1;
  while (c < kMaxCluster && (D / (c * 2)) >= 512 && (D % (c * 2 * 4)) == 0) c *= 2;
  const int e = knob(KNOB_TUPLE_CLUSTER);
  if (e != kKnobUnset) {
    if ((e == 1 || e == 2 || e == 4 || e == 8) && D % (4 * e) == 0) c = e;
  }
  pl->cluster = c;
  pl->Ds = D / c;
  const size_t budget = 110 * 1024;
  int dc = pl->Ds;
  if (anchor_smem_bytes(pl->sg, dc) > budget) {
    dc = 512;
    while (dc > 4 && (pl->Ds % dc) != 0) dc >>= 1;
    if (pl->Ds % dc) return SCL_ERR_BAD_SHAPE;
  }
  pl->Dc = dc;
  pl->smem = anchor_smem_bytes(pl->sg, dc);
  return SCL_OK;
}

template <int SG>
static int anchor_launch(const TupPlan& pl, const float* emb, int T, int S, int D, const AnchorArgs& a, float* demb,
                         float* loss, unsigned int* ws, cudaStream_t stream) {
  auto kern = anchor_tuple_kernel<SG>;
  static SmemAttrCache configured;            // per device
  int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), pl.smem, &configured);
  if (rc_attr) return rc_attr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(T) * pl.cluster);
  cfg.blockDim = dim3(kTupThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SCL_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, emb, T, S, D, pl.Ds, pl.Dc, a, demb, loss, ws));
  return SCL_OK;
}

static int anchor_run(const float* emb, int T, int S, int D, const AnchorArgs& a, float* loss, float* demb,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!emb || !loss || !workspace || T < 1) return SCL_ERR_BAD_ARG;
  if (!aligned16(emb) || (demb && !aligned16(demb)) || !aligned16(workspace)) return SCL_ERR_ALIGN;
  int rc = check_device();
  if (rc) return rc;
  TupPlan pl;
  rc = anchor_plan(S, D, &pl);
  if (rc) return rc;
  if (workspace_bytes < carve_bytes(4 + size_t(T), sizeof(float))) return SCL_ERR_WORKSPACE;
  unsigned int* ws = static_cast<unsigned int*>(workspace);
  SCL_CUDA_TRY(cudaMemsetAsync(ws, 0, 16, stream));
  switch (pl.sg) {
    case 25: return anchor_launch<25>(pl, emb, T, S, D, a, demb, loss, ws, stream);
    case 30: return anchor_launch<30>(pl, emb, T, S, D, a, demb, loss, ws, stream);
    default: return anchor_launch<35>(pl, emb, T, S, D, a, demb, loss, ws, stream);
  }
}

// ---------------------------------------------------------------------------------------------
// D1: out[t,i,j] = r_i - 2 x_i.x_j + r_j   (model/losses.py:656-661).  One CTA per tuple, one warp per pair.
__global__ void __launch_bounds__(256) pairwise_sqdist_kernel(const float* __restrict__ x, int n, int D,
                                                              float* __restrict__ out) {
  const int t = blockIdx.x;
  const float* X = x + size_t(t) * n * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int pr = warp; pr < n * n; pr += nwarps) {
    const int i = pr / n, j = pr - i * n;
    if (j < i) continue;
    float ri = 0.0f, rj = 0.0f, dot = 0.0f;
    if ((D & 3) == 0) {
      for (int c4 = lane; c4 < (D >> 2); c4 += 32) {
        const float4 a = *reinterpret_cast<const float4*>(X + size_t(i) * D + 4 * c4);
        const float4 b = *reinterpret_cast<const float4*>(X + size_t(j) * D + 4 * c4);
        ri = fmaf(a.x, a.x, ri); ri = fmaf(a.y, a.y, ri); ri = fmaf(a.z, a.z, ri); ri = fmaf(a.w, a.w, ri);
        rj = fmaf(b.x, b.x, rj); rj = fmaf(b.y, b.y, rj); rj = fmaf(b.z, b.z, rj); rj = fmaf(b.w, b.w, rj);
        dot = fmaf(a.x, b.x, dot); dot = fmaf(a.y, b.y, dot); dot = fmaf(a.z, b.z, dot); dot = fmaf(a.w, b.w, dot);
      }
    } else {
      for (int c = lane; c < D; c += 32) {
        float a = X[size_t(i) * D + c], b = X[size_t(j) * D + c];
        ri = fmaf(a, a, ri); rj = fmaf(b, b, rj); dot = fmaf(a, b, dot);
      }
    }
    ri = warp_sum(ri); rj = warp_sum(rj); dot = warp_sum(dot);
    if (lane == 0) {
      float v = ri - 2.0f * dot + rj;
      out[(size_t(t) * n + i) * n + j] = v;
      out[(size_t(t) * n + j) * n + i] = v;
    }
  }
}

}  // namespace scl

extern "C" int scl_tuple_loss_workspace_bytes(int T, int P, int N, int D, size_t* bytes) {
  (void)P; (void)N; (void)D;
  if (!bytes || T < 1) return SCL_ERR_BAD_ARG;
  *bytes = scl::carve_bytes(4 + size_t(T), sizeof(float));
  return SCL_OK;
}

extern "C" int scl_tuple_loss_fwd_bwd(const float* emb, int T, int P, int N, int D, const float* sq_d_dists,
                                      const scl_tuple_params* p, float* loss, float* demb, void* workspace,
                                      size_t workspace_bytes, scl_stream_t stream) {
  if (!p) return SCL_ERR_BAD_ARG;
  if (p->kind < SCL_TRIPLET || p->kind > SCL_DISTANCE_LAZY_QUADRUPLET) return SCL_ERR_BAD_ARG;
  if (p->dist_term < SCL_DIST_NONE || p->dist_term > SCL_DIST_HUBER) return SCL_ERR_BAD_ARG;
  if (p->dist_term != SCL_DIST_NONE && !sq_d_dists) return SCL_ERR_BAD_ARG;
  const bool dquad = p->kind == SCL_DISTANCE_QUADRUPLET || p->kind == SCL_DISTANCE_LAZY_QUADRUPLET;
  if (dquad && p->dist_term == SCL_DIST_NONE) return SCL_ERR_BAD_ARG;      // its second hinge is built on the distance term
  if (P < 1 || N < 1 || P > 32 || N > 32) return SCL_ERR_BAD_SHAPE;
  const bool quad = dquad || p->kind == SCL_QUADRUPLET || p->kind == SCL_LAZY_QUADRUPLET || p->kind == SCL_EVIL_QUADRUPLET;
  scl::AnchorArgs a = {};
  a.mode = scl::kModeTriplet;
  a.P = P; a.N = N; a.has_other = quad ? 1 : 0;
  a.tp = *p;
  a.sq_d_dists = sq_d_dists;
  const int S = 1 + P + N + (quad ? 1 : 0);
  return scl::anchor_run(emb, T, S, D, a, loss, demb, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int scl_logratio_fwd_bwd(const float* emb, int T, int P, int N, int D, const float* sq_pos,
                                    const float* sq_neg, int strict_reference, float* loss, float* demb,
                                    void* workspace, size_t workspace_bytes, scl_stream_t stream) {
  if (!sq_pos || !sq_neg) return SCL_ERR_BAD_ARG;
  if (P < 1 || N < 1 || P > 32 || N > 32) return SCL_ERR_BAD_SHAPE;
  if (strict_reference && P != N) return SCL_ERR_BAD_SHAPE;   // the reference's broadcast only exists for P == N
  scl::AnchorArgs a = {};
  a.mode = scl::kModeLogratio;
  a.P = P; a.N = N; a.has_other = 0;
  a.sq_pos = sq_pos; a.sq_neg = sq_neg;
  a.strict_reference = strict_reference;
  return scl::anchor_run(emb, T, 1 + P + N, D, a, loss, demb, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int scl_pairwise_sqdist(const float* x, int T, int n, int D, float* out, scl_stream_t stream) {
  if (!x || !out || T < 1 || n < 1 || D < 1) return SCL_ERR_BAD_ARG;
  if ((D & 3) == 0 && !scl::aligned16(x)) return SCL_ERR_ALIGN;
  int rc = scl::check_device();
  if (rc) return rc;
  scl::pairwise_sqdist_kernel<<<T, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, D, out);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}
