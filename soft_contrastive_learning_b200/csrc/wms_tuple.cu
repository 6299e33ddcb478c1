// wms_tuple.cu -- W1 tuple mode, C-ABI entry point and the CHUNKED kernel for tuples whose column slice does not
// fit in shared memory (D = 32768); everything else runs wms_tuple_resident.cu.  Fused forward + analytic backward
// of the weighted multi-similarity loss.
//
// Replaces wms_loss (/root/reference/model/losses.py:5-60, call train/train.py:852) and the TF autodiff of
// it (train.py:874-878) for T independent tuples of S <= 32 descriptors.
//
// One thread-block CLUSTER per tuple.  CTA c of the cluster owns the descriptor columns
// [c*Ds, (c+1)*Ds) of all S rows:
//   1. the [S x Ds] slice is brought into shared memory once (cp.async, 16 B per request);
//   2. the partial Gram matrix of the slice is accumulated on the FP32 pipes with TSxTS register tiles
//      (the 5x5 tile grid is symmetric, 15 tiles x 2 column groups fill a warp);
//   3. partial Grams are exchanged through distributed shared memory (st.shared::cluster push + one
//      cluster barrier) so every CTA holds the full S x S Gram;
//   4. every CTA redundantly evaluates norms, soft GPS masks, mining thresholds, the log-sum-exp weights
//      (ms_row.cuh) and folds the l2-normalisation Jacobian into one S x S matrix M;
//   5. d loss / d emb for the slice is M * E straight out of the still-resident shared-memory slice.
// HBM traffic is therefore the algorithmic minimum: emb read once, demb written once, dist read once.
// Slices that do not fit in shared memory (D = 32768) are streamed in chunks and re-read for step 5.
#include <atomic>
#include <cstdlib>

#include "ms_row.cuh"
#include "tuple_common.cuh"

namespace scl {

constexpr int kWmsThreads = kTupThreads;
constexpr int kWmsWarps = kTupWarps;
constexpr int kTileGrid = 5;                                   // 5 x 5 grid of TS x TS tiles
constexpr int kNumTiles = kTileGrid * (kTileGrid + 1) / 2;     // 15 symmetric tiles
// internal values of scl_ms_params::sumfunction that route the Gram skeleton to pairwise_distance_loss
constexpr int kSumPairwiseSquared = 100, kSumPairwiseHuber = 101;

template <int TS>
struct WmsSmem {
  static constexpr int SG = kTileGrid * TS;          // padded row count
  static constexpr int NP = kNumTiles * TS * TS;     // partial-Gram values in tile order
  static constexpr int HR = (SG + 1) / 2;            // output rows per half in the backward
  static constexpr int HRP = (HR + 3) / 4 * 4;
  // float offsets
  static __host__ __device__ size_t floats(int chunk_cols) {
    size_t n = 0;
    n += size_t(SG) * (chunk_cols + 4);   // Es
    n += al4(size_t(kWmsWarps) * NP);     // red
    n += al4(size_t(kMaxCluster) * NP);   // slots
    n += al4(size_t(SG) * (SG + 1));      // Gf
    n += size_t(SG) * (2 * HRP);          // Mt
    n += al4(4 * SG);                     // invn, flag, rowloss, c
    return n;
  }
};

__constant__ unsigned char c_tile_a[kNumTiles] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4};
__constant__ unsigned char c_tile_b[kNumTiles] = {0, 1, 2, 3, 4, 1, 2, 3, 4, 2, 3, 4, 3, 4, 4};

template <int TS>
__global__ void __launch_bounds__(kWmsThreads, (TS <= 6 ? 2 : 1)) wms_tuple_chunked_kernel(
    const float* __restrict__ emb, const float* __restrict__ dist, int T, int S, int D, int Ds, int Dc,
    scl_ms_params p, float* __restrict__ per_tuple, float* __restrict__ demb, uint32_t* __restrict__ kept,
    float* __restrict__ loss_out, unsigned int* __restrict__ done_counter) {
  using L = WmsSmem<TS>;
  constexpr int SG = L::SG, NP = L::NP, HRP = L::HRP;
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int crank = cluster.block_rank();
  const int t = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pitch = Dc + 4;
  const int ncols4 = Dc >> 2;       // float4 columns per chunk
  const int nchunks = Ds / Dc;

  float* Es = smem;
  float* red = Es + size_t(SG) * pitch;
  float* slots = red + al4(kWmsWarps * NP);
  float* Gf = slots + al4(kMaxCluster * NP);
  float* Mt = Gf + al4(SG * (SG + 1));   // 16-byte aligned: read as float4
  float* invn = Mt + SG * 2 * HRP;
  float* nflag = invn + SG;
  float* rowloss = nflag + SG;
  float* cvec = rowloss + SG;

  const float* E_t = emb + (size_t(t) * S) * D + size_t(crank) * Ds;
  // every CTA of the cluster must be resident before anyone writes into a peer's shared memory:
  // arrive now, wait just before the push
  if (C > 1) cluster_arrive();

  // zero the padding rows once (rows >= S contribute nothing to any tile)
  for (int i = S * pitch + tid; i < SG * pitch; i += kWmsThreads) Es[i] = 0.0f;

  auto load_chunk = [&](int ch) { tup_load_chunk(Es, E_t + size_t(ch) * Dc, S, D, Dc); };

  // ---------------- forward: partial Gram over this CTA's slice ----------------
  const int tl = lane % kNumTiles;
  const int dg = lane / kNumTiles;                 // 0,1 (lanes 30,31 -> 2: idle)
  const int ta = c_tile_a[tl], tb = c_tile_b[tl];
  const int dgid = warp * 2 + dg;                  // 16 column groups per CTA
  float acc[TS][TS];
#pragma unroll
  for (int r = 0; r < TS; ++r)
#pragma unroll
    for (int q = 0; q < TS; ++q) acc[r][q] = 0.0f;

  for (int ch = 0; ch < nchunks; ++ch) {
    load_chunk(ch);
    if (dg < 2) {
      for (int c4 = dgid; c4 < ncols4; c4 += 2 * kWmsWarps) {
        float4 x[TS], y[TS];
#pragma unroll
        for (int r = 0; r < TS; ++r) {
          x[r] = *reinterpret_cast<const float4*>(Es + (ta + kTileGrid * r) * pitch + 4 * c4);
          y[r] = *reinterpret_cast<const float4*>(Es + (tb + kTileGrid * r) * pitch + 4 * c4);
        }
#pragma unroll
        for (int r = 0; r < TS; ++r)
#pragma unroll
          for (int q = 0; q < TS; ++q) {
            acc[r][q] = fmaf(x[r].x, y[q].x, acc[r][q]);
            acc[r][q] = fmaf(x[r].y, y[q].y, acc[r][q]);
            acc[r][q] = fmaf(x[r].z, y[q].z, acc[r][q]);
            acc[r][q] = fmaf(x[r].w, y[q].w, acc[r][q]);
          }
      }
    }
    if (nchunks > 1) __syncthreads();              // chunk buffer is about to be overwritten
  }

  // fold the two column groups of a warp, then the warps of the CTA
#pragma unroll
  for (int r = 0; r < TS; ++r)
#pragma unroll
    for (int q = 0; q < TS; ++q) {
      float o = __shfl_down_sync(0xffffffffu, acc[r][q], kNumTiles);
      if (lane < kNumTiles) red[warp * NP + tl * TS * TS + r * TS + q] = acc[r][q] + o;
    }
  __syncthreads();
  if (C > 1) cluster_wait();
  for (int k = tid; k < NP; k += kWmsThreads) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < kWmsWarps; ++w) s += red[w * NP + k];
    if (C == 1) {
      slots[k] = s;
    } else {
      // push this CTA's partial into slot [crank] of every CTA of the cluster (distributed shared memory)
      for (int peer = 0; peer < C; ++peer) {
        float* remote = cluster.map_shared_rank(slots, peer);
        remote[crank * NP + k] = s;
      }
    }
  }
  if (C > 1) cluster.sync(); else __syncthreads();
  for (int k = tid; k < NP; k += kWmsThreads) {
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += slots[c * NP + k];
    int tile = k / (TS * TS), rq = k - tile * TS * TS;
    int r = rq / TS, q = rq - r * TS;
    int i = c_tile_a[tile] + kTileGrid * r, j = c_tile_b[tile] + kTileGrid * q;
    Gf[i * (SG + 1) + j] = s;
    if (c_tile_a[tile] != c_tile_b[tile]) Gf[j * (SG + 1) + i] = s;
  }
  __syncthreads();

  float* Mraw = red + 2 * SG * SG;
  if (p.sumfunction >= kSumPairwiseSquared) {
    // ---------------- pairwise_distance_loss (model/losses.py:627-646) on the same Gram ----------------
    // F_ij = r_i - 2 x_i.x_j + r_j (losses.py:656-661), e = F/f_max - d/d_max, squared or Huber(delta = 1), mean over
    // all n*n pairs;  A_ij = dL_t/dF_ij,  dX = M X with M = 2 (diag(rowsum(A + A^T)) - (A + A^T)).
    // p.alpha carries d_max_squared, p.beta f_max_squared (internal call, see scl_pairwise_distance_loss_fwd_bwd).
    const float* dist_t = dist + size_t(t) * S * S;
    const float inv_nn = 1.0f / float(S * S);
    for (int i = warp; i < S; i += kWmsWarps) {
      float el = 0.0f, g = 0.0f;
      if (lane < S) {
        const float F = Gf[i * (SG + 1) + i] - 2.0f * Gf[i * (SG + 1) + lane] + Gf[lane * (SG + 1) + lane];
        const float e = F / p.beta - dist_t[i * S + lane] / p.alpha;
        if (p.sumfunction == kSumPairwiseHuber) {
          const float ae = fabsf(e), q = fminf(ae, 1.0f);
          el = 0.5f * q * q + (ae - q);
          g = (ae <= 1.0f) ? e : (e > 0.0f ? 1.0f : -1.0f);
        } else {
          el = e * e;
          g = 2.0f * e;
        }
        red[i * SG + lane] = g * inv_nn / p.beta;
      }
      el = warp_sum(el);
      if (lane == 0) rowloss[i] = el * inv_nn;
    }
    __syncthreads();
    for (int i = warp; i < S; i += kWmsWarps) {
      const float w = lane < S ? red[i * SG + lane] + red[lane * SG + i] : 0.0f;
      const float rs = warp_sum(w);
      if (lane == 0) cvec[i] = rs;
    }
    __syncthreads();
    for (int k = tid; k < S * S; k += kWmsThreads) {
      const int i = k / S, j = k - i * S;
      const float w = red[i * SG + j] + red[j * SG + i];
      Mraw[i * SG + j] = (i == j) ? 2.0f * (cvec[i] - w) : -2.0f * w;
    }
    __syncthreads();
  } else {
  // ---------------- weights: one warp per anchor row ----------------
    if (tid < SG) {
      float n2 = tid < S ? Gf[tid * (SG + 1) + tid] : 1.0f;
      // tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))  (losses.py:7)
      invn[tid] = rsqrtf(fmaxf(n2, 1e-12f));
      nflag[tid] = n2 >= 1e-12f ? 1.0f : 0.0f;     // below the clamp the normalisation is a pure scale
    }
    __syncthreads();
    const float* dist_t = dist + size_t(t) * S * S;
    const float invS = 1.0f / float(S);
    for (int i = warp; i < S; i += kWmsWarps) {
      const int j = lane;
      const bool valid = j < S;
      float raw = 0.0f, s = 0.0f, wp = 0.0f, wn = 0.0f;
      if (valid) {
        raw = Gf[i * (SG + 1) + j] * invn[i] * invn[j];
        s = fmaxf(raw, 0.0f);                                              // losses.py:26
        wms_masks(dist_t[i * S + j], p.d_alpha, p.d_beta, p.wfunction, wp, wn);
        if (i == j) wp -= 1.0f;                                            // losses.py:22
      }
      MsRowStats st;
      st.maxv = warp_max(valid ? s * wn : -INFINITY);
      st.tmp = warp_max(valid ? s * wp : -INFINITY);
      st.minv = warp_min(valid ? (s - st.tmp) * wp : INFINITY) + st.tmp;
      bool kp = false, kn = false;
      float ep = 0.0f, en = 0.0f;
      if (valid) ms_elem(s, wp, wn, st, p, kp, kn, ep, en);
      float A = warp_sum(ep), B = warp_sum(en);
      float g = 0.0f;
      if (valid) {
        g = ms_elem_grad(wp, wn, kp, kn, ep, en, A, B, p) * invS;
        if (!(raw >= 0.0f)) g = 0.0f;                                      // tf.maximum passes gradient when x >= 0
      }
      // Gf now holds similarities in the upper use; stash dL/ds into `red` (free again) as Gw[i][j]
      if (valid) red[i * SG + j] = g;
      if (valid) red[SG * SG + i * SG + j] = raw;
      if (lane == 0) rowloss[i] = ms_row_loss(A, B, p) * invS;
      if (kept != nullptr && crank == 0) {
        unsigned mp = __ballot_sync(0xffffffffu, kp), mn = __ballot_sync(0xffffffffu, kn);
        if (lane == 0) {
          kept[(size_t(t) * S + i) * 2 + 0] = mp;
          kept[(size_t(t) * S + i) * 2 + 1] = mn;
        }
      }
    }
    __syncthreads();
    // M = diag(invn) (W - diag(c)) diag(invn), W = Gw + Gw^T, c_i = sum_j W_ij s_ij(raw)  (projection of l2norm)
    const float* Gw = red;
    const float* Sraw = red + SG * SG;
    for (int i = warp; i < S; i += kWmsWarps) {
      float w = 0.0f, part = 0.0f;
      if (lane < S) {
        w = Gw[i * SG + lane] + Gw[lane * SG + i];
        part = w * Sraw[i * SG + lane];
      }
      float c = warp_sum(part) * nflag[i];
      if (lane == 0) cvec[i] = c;
    }
    __syncthreads();
    // overwrite Gw in place with M (Gw/Sraw are not needed afterwards), then store it transposed for step 5
    for (int k = tid; k < S * S; k += kWmsThreads) {
      int i = k / S, j = k - i * S;
      float w = Gw[i * SG + j] + Gw[j * SG + i];
      if (i == j) w -= cvec[i];
      Mraw[i * SG + j] = invn[i] * w * invn[j];
    }
    __syncthreads();
  }
  tup_store_Mt<SG>(Mt, Mraw, SG, S);

  // ---------------- loss ----------------
  if (crank == 0 && warp == 0) {
    float v = lane < S ? rowloss[lane] : 0.0f;
    v = warp_sum(v);
    if (lane == 0 && per_tuple != nullptr) per_tuple[t] = v;
    tup_finish_loss(done_counter, t, T, v, loss_out, lane);
  }

  // ---------------- backward: demb slice = (1/T) * M * E_slice ----------------
  if (demb != nullptr) {
    float* dE_t = demb + (size_t(t) * S) * D + size_t(crank) * Ds;
    for (int ch = 0; ch < nchunks; ++ch) {
      if (nchunks > 1) load_chunk(ch);
      tup_bwd_chunk<SG>(Es, Mt, dE_t + size_t(ch) * Dc, S, D, Dc, 1.0f / float(T));
      if (nchunks > 1) __syncthreads();
    }
  }
  // a CTA must not exit while peers may still push into its shared memory: the only remote writes happen
  // before the cluster.sync() above, so no trailing barrier is needed.
}

// ---------------------------------------------------------------------------------------------
struct WmsPlan {
  int ts;        // 5, 6 or 7
  int cluster;   // 1,2,4,8
  int Ds, Dc;
  size_t smem;
};

static size_t wms_smem_bytes(int ts, int dc) {
  switch (ts) {
    case 5: return WmsSmem<5>::floats(dc) * sizeof(float);
    case 6: return WmsSmem<6>::floats(dc) * sizeof(float);
    default: return WmsSmem<7>::floats(dc) * sizeof(float);
  }
}

static int wms_plan(int S, int D, WmsPlan* pl) {
  if (S < 2 || S > 32 || D < 4 || (D & 3)) return SCL_ERR_BAD_SHAPE;
  pl->ts = S <= 25 ? 5 : (S <= 30 ? 6 : 7);
  // cluster size: keep the per-CTA slice around 512 columns (two CTAs per SM stay resident)
  int c = 1;
  while (c < kMaxCluster && (D / (c * 2)) >= 512 && (D % (c * 2 * 4)) == 0) c *= 2;
  const int e = knob(KNOB_WMS_CLUSTER);
  if (e != kKnobUnset) {
    if ((e == 1 || e == 2 || e == 4 || e == 8) && D % (4 * e) == 0) c = e;
  }
  pl->cluster = c;
  pl->Ds = D / c;
  // chunk: whole slice if it fits in ~100 KB of shared memory, else 512-column chunks
  const size_t budget = 110 * 1024;
  int dc = pl->Ds;
  if (wms_smem_bytes(pl->ts, dc) > budget) {
    dc = 512;
    while (dc > 4 && (pl->Ds % dc) != 0) dc >>= 1;
    if (pl->Ds % dc) return SCL_ERR_BAD_SHAPE;
  }
  pl->Dc = dc;
  pl->smem = wms_smem_bytes(pl->ts, dc);
  if (pl->smem > 227 * 1024) return SCL_ERR_BAD_SHAPE;
  return SCL_OK;
}

template <int TS>
static int wms_launch_chunked(const WmsPlan& pl, const float* emb, const float* dist, int T, int S, int D,
                      const scl_ms_params& p, float* per_tuple, float* demb, uint32_t* kept, float* loss,
                      unsigned int* counter, cudaStream_t stream) {
  auto kern = wms_tuple_chunked_kernel<TS>;
  static SmemAttrCache configured;            // per device; set only when it has to grow
  int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), pl.smem, &configured);
  if (rc_attr) return rc_attr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(T) * pl.cluster);
  cfg.blockDim = dim3(kWmsThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SCL_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, emb, dist, T, S, D, pl.Ds, pl.Dc, p, per_tuple, demb, kept, loss,
                                  counter));
  return SCL_OK;
}

int wms_stream_launch(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params& p, float* loss,
                      float* per_tuple, float* demb, uint32_t* kept, unsigned int* counter, cudaStream_t stream);
int wms_resident_launch(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params& p, float* loss,
                        float* per_tuple, float* demb, uint32_t* kept, unsigned int* counter, cudaStream_t stream);

}  // namespace scl

extern "C" int scl_pairwise_distance_loss_fwd_bwd(const float* emb, const float* pairwise_sq_d, int T, int n, int D,
                                                  float d_max_squared, float f_max_squared, int huber, float* loss,
                                                  float* demb, void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!emb || !pairwise_sq_d || !loss || !workspace) return SCL_ERR_BAD_ARG;
  if (T < 1 || !(d_max_squared > 0.0f) || !(f_max_squared > 0.0f)) return SCL_ERR_BAD_ARG;
  if (!scl::aligned16(emb) || (demb && !scl::aligned16(demb)) || !scl::aligned16(workspace)) return SCL_ERR_ALIGN;
  int rc = scl::check_device();
  if (rc) return rc;
  scl::WmsPlan pl;
  rc = scl::wms_plan(n, D, &pl);
  if (rc) return rc;
  size_t need = 0;
  scl_wms_tuple_workspace_bytes(T, n, D, &need);
  if (workspace_bytes < need) return SCL_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  unsigned int* counter = static_cast<unsigned int*>(workspace);
  SCL_CUDA_TRY(cudaMemsetAsync(counter, 0, 16, stream));
  scl_ms_params p = {};
  p.alpha = d_max_squared;
  p.beta = f_max_squared;
  p.sumfunction = huber ? scl::kSumPairwiseHuber : scl::kSumPairwiseSquared;
  switch (pl.ts) {
    case 5: return scl::wms_launch_chunked<5>(pl, emb, pairwise_sq_d, T, n, D, p, nullptr, demb, nullptr, loss, counter, stream);
    case 6: return scl::wms_launch_chunked<6>(pl, emb, pairwise_sq_d, T, n, D, p, nullptr, demb, nullptr, loss, counter, stream);
    default: return scl::wms_launch_chunked<7>(pl, emb, pairwise_sq_d, T, n, D, p, nullptr, demb, nullptr, loss, counter, stream);
  }
}

extern "C" int scl_wms_tuple_workspace_bytes(int T, int S, int D, size_t* bytes) {
  if (!bytes || T < 1) return SCL_ERR_BAD_ARG;
  scl::WmsPlan pl;
  int rc = scl::wms_plan(S, D, &pl);
  if (rc) return rc;
  *bytes = scl::carve_bytes(4 + size_t(T), sizeof(float));
  return SCL_OK;
}

extern "C" int scl_wms_tuple_fwd_bwd(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params* p,
                                     float* loss, float* per_tuple, float* demb, uint32_t* kept, void* workspace,
                                     size_t workspace_bytes, scl_stream_t stream_) {
  if (!emb || !dist || !p || !loss || !workspace || T < 1) return SCL_ERR_BAD_ARG;
  if (!scl::aligned16(emb) || (demb && !scl::aligned16(demb)) || !scl::aligned16(workspace)) return SCL_ERR_ALIGN;
  int rc = scl::check_device();
  if (rc) return rc;
  scl::WmsPlan pl;
  rc = scl::wms_plan(S, D, &pl);
  if (rc) return rc;
  size_t need = 0;
  scl_wms_tuple_workspace_bytes(T, S, D, &need);
  if (workspace_bytes < need) return SCL_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  unsigned int* counter = static_cast<unsigned int*>(workspace);
  SCL_CUDA_TRY(cudaMemsetAsync(counter, 0, 16, stream));
  // large batches: streaming kernel (one tuple per CTA); small ones: a cluster per tuple (resident slice if it fits)
  rc = scl::wms_stream_launch(emb, dist, T, S, D, *p, loss, per_tuple, demb, kept, counter, stream);
  if (rc != SCL_ERR_UNSUPPORTED) return rc;
  rc = scl::wms_resident_launch(emb, dist, T, S, D, *p, loss, per_tuple, demb, kept, counter, stream);
  if (rc != SCL_ERR_UNSUPPORTED) return rc;
  switch (pl.ts) {
    case 5: return scl::wms_launch_chunked<5>(pl, emb, dist, T, S, D, *p, per_tuple, demb, kept, loss, counter, stream);
    case 6: return scl::wms_launch_chunked<6>(pl, emb, dist, T, S, D, *p, per_tuple, demb, kept, loss, counter, stream);
    default: return scl::wms_launch_chunked<7>(pl, emb, dist, T, S, D, *p, per_tuple, demb, kept, loss, counter, stream);
  }
}
