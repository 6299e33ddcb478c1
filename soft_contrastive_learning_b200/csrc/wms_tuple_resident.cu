// wms_tuple_resident.cu -- W1 tuple mode, main path: fused forward + analytic backward of the weighted
// multi-similarity loss with the tuple's descriptors resident in shared memory.
//
// Replaces wms_loss (/root/reference/model/losses.py:5-60, call train/train.py:852) and the TF autodiff of it
// (train.py:874-878) for T independent tuples of S <= 32 descriptors.
//
// One thread-block CLUSTER of C CTAs per tuple; CTA c owns the descriptor columns [c*Ds, (c+1)*Ds) of all S rows.
//   1. warp 0 brings the [S x Ds] slice in with bulk async copies (cp.async.bulk, one per row and column chunk,
//      completion on one mbarrier per chunk), so the Gram accumulation starts as soon as the first chunk has landed;
//      meanwhile every thread evaluates the GPS soft masks of "its" pairs (losses.py:11-19) into registers;
//   2. partial Gram of the slice on the FP32 pipes with packed FFMA2 (fma.rn.f32x2): the symmetric 5x5 grid of
//      TSxTS register tiles, 15 tiles x 2 column groups per warp, even/odd-column partial sums in one register pair;
//   3. partial Grams are exchanged through distributed shared memory (st.shared::cluster push + one cluster
//      barrier); every CTA sums them in the same order, so all CTAs of a tuple take bit-identical mining decisions;
//   4. every CTA redundantly evaluates mining thresholds and log-sum-exp weights (ms_row.cuh) and folds the
//      l2-normalisation Jacobian and the 1/T batch mean into one S x S matrix M;
//   5. d loss / d emb for the slice is M * E straight out of the resident slice, again with FFMA2 (the M factor is a
//      broadcast scalar operand), written with 512-byte-per-row coalesced streaming stores.
// HBM traffic is the algorithmic minimum: emb read once, demb written once, dist read once per CTA (L2 hits).
// Shapes whose slice does not fit (e.g. D = 32768) go to the chunked kernel in wms_tuple.cu.
#include <atomic>
#include <cstdlib>

#include "ms_row.cuh"
#include "tc_common.cuh"
#include "tuple_common.cuh"

namespace scl {

using namespace tc;

constexpr int kRThreads = 256;
constexpr int kRWarps = kRThreads / 32;
constexpr int kRGrid = 5;                                  // 5 x 5 grid of TS x TS tiles
constexpr int kRTiles = kRGrid * (kRGrid + 1) / 2;         // 15 symmetric tiles
constexpr int kRChunks = 4;                                // column chunks (one mbarrier each)
constexpr int kRRed = 4;                                   // cross-warp reduction buffers (two rounds)

template <int TS>
struct RSmem {
  static constexpr int SG = kRGrid * TS;
  static constexpr int NP = kRTiles * TS * TS;
  static constexpr int HR = (SG + 1) / 2;
  static constexpr int HRP = (HR + 3) / 4 * 4;
  static constexpr int MT = SG * 2 * HRP;                              // transposed M (overlays the slots)
  static constexpr int RED = (kRRed * NP > 2 * SG * SG) ? kRRed * NP : 2 * SG * SG;   // reduction, then Gw and Sraw
  static __host__ __device__ size_t slots_floats(int C) {
    size_t a = size_t(C) * NP, b = MT;
    return al4(a > b ? a : b);
  }
  static __host__ __device__ size_t bytes(int Ds, int C) {
    size_t f = size_t(SG) * (Ds + 4) + al4(RED) + slots_floats(C) + al4(4 * SG);
    return f * sizeof(float) + kRChunks * sizeof(uint64_t);
  }
};

__constant__ unsigned char c_rtile_a[kRTiles] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4};
__constant__ unsigned char c_rtile_b[kRTiles] = {0, 1, 2, 3, 4, 1, 2, 3, 4, 2, 3, 4, 3, 4, 4};

// d += a * b on both halves of a register pair (SASS FFMA2); a scalar operand is passed as make_float2(m, m) and
// folded by ptxas into the instruction's broadcast form.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}

// 1-D bulk async copy global -> shared of this CTA, completion (bytes) on an mbarrier of this CTA
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int TS>
__global__ void __launch_bounds__(kRThreads, (TS == 5 ? 2 : 1)) wms_tuple_kernel(
    const float* __restrict__ emb, const float* __restrict__ dist, int T, int S, int D, int Ds, scl_ms_params p,
    float* __restrict__ per_tuple, float* __restrict__ demb, uint32_t* __restrict__ kept, float* __restrict__ loss_out,
    unsigned int* __restrict__ done_counter) {
  using L = RSmem<TS>;
  constexpr int SG = L::SG, NP = L::NP, HR = L::HR, HRP = L::HRP;
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int crank = cluster.block_rank();
  const int t = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pitch = Ds + 4;
  const int ncols4 = Ds >> 2;

  float* Es = smem;
  float* red = Es + size_t(SG) * pitch;          // cross-warp reduction, later Gw and Sraw
  float* slots = red + al4(L::RED);              // cluster exchange, later Mt
  float* misc = slots + L::slots_floats(C);
  float* rowloss = misc;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + al4(4 * SG));

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < kRChunks; ++k) mbar_init(&bars[k], 1);
    fence_barrier_init();
  }
  // zero the padding rows once (rows >= S contribute nothing to any tile)
  for (int i = S * pitch + tid; i < SG * pitch; i += kRThreads) Es[i] = 0.0f;
  __syncthreads();
  // every CTA of the cluster must be resident before anyone writes into a peer's shared memory: arrive now,
  // wait just before the push
  if (C > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  // ---------------- 1. slice -> shared memory ----------------
  const float* E_t = emb + (size_t(t) * S) * D + size_t(crank) * Ds;
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < kRChunks; ++k) {
      const int q0 = (k * ncols4) / kRChunks, q1 = ((k + 1) * ncols4) / kRChunks;
      const uint32_t bytes = uint32_t(q1 - q0) * 16u;
      if (lane == 0) {
        if (bytes) mbar_arrive_expect_tx(&bars[k], bytes * uint32_t(S)); else mbar_arrive(&bars[k]);
      }
      __syncwarp();
      if (lane < S && bytes) bulk_load_1d(Es + lane * pitch + 4 * q0, E_t + size_t(lane) * D + 4 * q0, bytes, &bars[k]);
    }
  }

  // GPS soft masks of this thread's pairs: row i = warp + 8k, column j = lane  (losses.py:11-22)
  const float* dist_t = dist + size_t(t) * S * S;
  float wpv[4], wnv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = warp + kRWarps * k;
    wpv[k] = 0.0f;
    wnv[k] = 0.0f;
    if (i < S && lane < S) {
      wms_masks(__ldg(dist_t + i * S + lane), p.d_alpha, p.d_beta, p.wfunction, wpv[k], wnv[k]);
      if (i == lane) wpv[k] -= 1.0f;                                       // losses.py:22
    }
  }

  // ---------------- 2. partial Gram over this CTA's slice ----------------
  const int tl = lane % kRTiles;
  const int dg = lane / kRTiles;                   // 0,1 (lanes 30,31 -> 2: idle)
  const int ta = c_rtile_a[tl], tb = c_rtile_b[tl];
  const int dgid = warp * 2 + dg;                  // 16 column groups per CTA
  float2 acc[TS][TS];
#pragma unroll
  for (int r = 0; r < TS; ++r)
#pragma unroll
    for (int q = 0; q < TS; ++q) acc[r][q] = make_float2(0.0f, 0.0f);

#pragma unroll 1
  for (int k = 0; k < kRChunks; ++k) {
    const int q0 = (k * ncols4) / kRChunks, q1 = ((k + 1) * ncols4) / kRChunks;
    mbar_wait(&bars[k], 0);
    if (dg < 2) {
#pragma unroll 1
      for (int c4 = q0 + dgid; c4 < q1; c4 += 2 * kRWarps) {
        float4 x[TS], y[TS];
#pragma unroll
        for (int r = 0; r < TS; ++r) {
          x[r] = *reinterpret_cast<const float4*>(Es + (ta + kRGrid * r) * pitch + 4 * c4);
          y[r] = *reinterpret_cast<const float4*>(Es + (tb + kRGrid * r) * pitch + 4 * c4);
        }
#pragma unroll
        for (int r = 0; r < TS; ++r)
#pragma unroll
          for (int q = 0; q < TS; ++q) {
            ffma2(acc[r][q], make_float2(x[r].x, x[r].y), make_float2(y[q].x, y[q].y));
            ffma2(acc[r][q], make_float2(x[r].z, x[r].w), make_float2(y[q].z, y[q].w));
          }
      }
    }
  }

  // fold even/odd columns, the two column groups of a warp, then the warps of the CTA (two rounds over kRRed buffers)
  float g[TS][TS];
#pragma unroll
  for (int r = 0; r < TS; ++r)
#pragma unroll
    for (int q = 0; q < TS; ++q) {
      const float v = acc[r][q].x + acc[r][q].y;
      g[r][q] = v + __shfl_down_sync(0xffffffffu, v, kRTiles);
    }
  if (warp >= kRRed && lane < kRTiles) {
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
      for (int q = 0; q < TS; ++q) red[(warp - kRRed) * NP + (r * TS + q) * kRTiles + tl] = g[r][q];
  }
  __syncthreads();
  if (warp < kRRed && lane < kRTiles) {
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
      for (int q = 0; q < TS; ++q) {
        const int a = warp * NP + (r * TS + q) * kRTiles + tl;
        red[a] += g[r][q];
      }
  }
  __syncthreads();
  if (C > 1) cluster_wait();
  for (int k = tid; k < NP; k += kRThreads) {
    const float s = (red[k] + red[NP + k]) + (red[2 * NP + k] + red[3 * NP + k]);
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

  // ---------------- 3. weights: warp w owns anchor rows w, w+8, w+16, w+24 (interleaved), lane = column j ----------------
  // G_ij is summed straight out of the exchange slots in a fixed order (identical in every CTA of the cluster).
  auto gram = [&](int i, int j) -> float {
    int a = i % kRGrid, r = i / kRGrid, b = j % kRGrid, q = j / kRGrid;
    if (a > b) { int x = a; a = b; b = x; x = r; r = q; q = x; }
    const int k = (r * TS + q) * kRTiles + (a * kRGrid - (a * (a - 1)) / 2 + (b - a));
    float v = 0.0f;
    for (int c = 0; c < C; ++c) v += slots[c * NP + k];
    return v;
  };
  float* Gw = red;                                // dL/ds_ij        [SG][SG]
  float* Sraw = red + SG * SG;                    // cosine similarity before the relu
  float* Mt = slots;                              // written after every warp is done with the slots
  const int jj = lane < S ? lane : 0;
  const float n2 = gram(jj, jj);
  // tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))  (losses.py:7); below the clamp it is a pure scale
  const float invn_j = rsqrtf(fmaxf(n2, 1e-12f));
  const float nflag_j = n2 >= 1e-12f ? 1.0f : 0.0f;
  const float invS = 1.0f / float(S);
  const bool jvalid = lane < S;
  float invn_i[4], raw[4], sv[4];
  bool rvalid[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = warp + kRWarps * k;
    rvalid[k] = i < S;
    const int ii = rvalid[k] ? i : 0;
    invn_i[k] = __shfl_sync(0xffffffffu, invn_j, ii);
    raw[k] = (rvalid[k] && jvalid) ? gram(ii, jj) * invn_i[k] * invn_j : 0.0f;
    sv[k] = fmaxf(raw[k], 0.0f);                                           // losses.py:26
  }
  MsRowStats st[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    st[k].maxv = jvalid ? sv[k] * wnv[k] : -INFINITY;
    st[k].tmp = jvalid ? sv[k] * wpv[k] : -INFINITY;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      st[k].maxv = fmaxf(st[k].maxv, __shfl_xor_sync(0xffffffffu, st[k].maxv, o));
      st[k].tmp = fmaxf(st[k].tmp, __shfl_xor_sync(0xffffffffu, st[k].tmp, o));
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) st[k].minv = jvalid ? (sv[k] - st[k].tmp) * wpv[k] : INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) st[k].minv = fminf(st[k].minv, __shfl_xor_sync(0xffffffffu, st[k].minv, o));
  bool kp[4], kn[4];
  float ep[4], en[4], A[4], B[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    st[k].minv += st[k].tmp;
    kp[k] = kn[k] = false;
    ep[k] = en[k] = 0.0f;
    if (jvalid && rvalid[k]) ms_elem(sv[k], wpv[k], wnv[k], st[k], p, kp[k], kn[k], ep[k], en[k]);
    A[k] = ep[k];
    B[k] = en[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      A[k] += __shfl_xor_sync(0xffffffffu, A[k], o);
      B[k] += __shfl_xor_sync(0xffffffffu, B[k], o);
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = warp + kRWarps * k;
    if (rvalid[k]) {
      if (jvalid) {
        float gw = ms_elem_grad(wpv[k], wnv[k], kp[k], kn[k], ep[k], en[k], A[k], B[k], p) * invS;
        if (!(raw[k] >= 0.0f)) gw = 0.0f;                                  // tf.maximum passes gradient when x >= 0
        Gw[i * SG + lane] = gw;
        Sraw[i * SG + lane] = raw[k];
      }
      if (lane == 0) rowloss[i] = ms_row_loss(A[k], B[k], p) * invS;
    }
    if (kept != nullptr && crank == 0) {
      const unsigned mp = __ballot_sync(0xffffffffu, kp[k]), mn = __ballot_sync(0xffffffffu, kn[k]);
      if (lane == 0 && rvalid[k]) {
        kept[(size_t(t) * S + i) * 2 + 0] = mp;
        kept[(size_t(t) * S + i) * 2 + 1] = mn;
      }
    }
  }
  __syncthreads();
  // M = (1/T) diag(invn) (W - diag(c)) diag(invn), W = Gw + Gw^T, c_i = sum_j W_ij s_ij(raw)  (projection of l2norm),
  // stored transposed and split in two row halves for the backward: Mt[j][h*HRP + r] = M[h*HR + r][j]
  const float invT = 1.0f / float(T);
  float wij[4], cpart[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = warp + kRWarps * k;
    wij[k] = (rvalid[k] && jvalid) ? Gw[i * SG + lane] + Gw[lane * SG + i] : 0.0f;
    cpart[k] = (rvalid[k] && jvalid) ? wij[k] * Sraw[i * SG + lane] : 0.0f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) cpart[k] += __shfl_xor_sync(0xffffffffu, cpart[k], o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = warp + kRWarps * k;
    const float c = cpart[k] * __shfl_sync(0xffffffffu, nflag_j, rvalid[k] ? i : 0);
    if (rvalid[k] && jvalid) {
      const float w = (i == lane) ? wij[k] - c : wij[k];
      const int h = i / HR, r = i - h * HR;
      Mt[lane * (2 * HRP) + h * HRP + r] = invn_i[k] * w * invn_j * invT;
    }
  }

  // ---------------- loss ----------------
  if (crank == 0 && warp == 0) {
    float v = lane < S ? rowloss[lane] : 0.0f;
    v = warp_sum(v);
    if (lane == 0 && per_tuple != nullptr) per_tuple[t] = v;
    tup_finish_loss(done_counter, t, T, v, loss_out, lane);
  }
  __syncthreads();

  // ---------------- 5. backward: demb slice = M * E_slice ----------------
  if (demb != nullptr) {
    float* dE_t = demb + (size_t(t) * S) * D + size_t(crank) * Ds;
    for (int item = tid; item < 2 * ncols4; item += kRThreads) {
      const int h = item / ncols4, c4 = item - h * ncols4;
      float2 o[HR][2];
#pragma unroll
      for (int r = 0; r < HR; ++r) o[r][0] = o[r][1] = make_float2(0.0f, 0.0f);
      const float* ecol = Es + 4 * c4;
      const float* mcol = Mt + h * HRP;
#pragma unroll 5
      for (int j = 0; j < S; ++j) {
        const float4 e = *reinterpret_cast<const float4*>(ecol + j * pitch);
        const float4* mrow = reinterpret_cast<const float4*>(mcol + j * (2 * HRP));
#pragma unroll
        for (int r4 = 0; r4 < HRP / 4; ++r4) {
          const float4 m = mrow[r4];
          const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r4 * 4 + u;
            if (r < HR) {
              ffma2(o[r][0], make_float2(mv[u], mv[u]), make_float2(e.x, e.y));
              ffma2(o[r][1], make_float2(mv[u], mv[u]), make_float2(e.z, e.w));
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        const int i = h * HR + r;
        if (i < S)
          stg_stream(reinterpret_cast<float4*>(dE_t + size_t(i) * D + 4 * c4),
                     make_float4(o[r][0].x, o[r][0].y, o[r][1].x, o[r][1].y));
      }
    }
  }
  // a CTA must not exit while peers may still push into its shared memory: the only remote writes happen
  // before the cluster.sync() above, so no trailing barrier is needed.
}

// ---------------------------------------------------------------------------------------------
static size_t resident_bytes(int ts, int Ds, int C) {
  switch (ts) {
    case 5: return RSmem<5>::bytes(Ds, C);
    case 6: return RSmem<6>::bytes(Ds, C);
    default: return RSmem<7>::bytes(Ds, C);
  }
}

struct ResidentPlan {
  int ts, cluster, Ds;
  size_t smem;
};

// Smallest cluster whose slice fits the per-CTA budget; small batches are spread over more CTAs.
static bool resident_plan(int T, int S, int D, ResidentPlan* pl) {
  if (S < 2 || S > 32 || D < 4 || (D & 3)) return false;
  const int ts = S <= 25 ? 5 : (S <= 30 ? 6 : 7);
  // TS = 5 runs two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2; the wider tiles need the whole register file
  const size_t budget = ts == 5 ? size_t(115712) : size_t(227 * 1024);
  int chosen = 0;
  for (int c = 1; c <= kMaxCluster; c *= 2) {
    if (D % (4 * c)) break;
    if (resident_bytes(ts, D / c, c) <= budget) { chosen = c; break; }
  }
  const int e = knob(KNOB_WMS_CLUSTER);
  if (e != kKnobUnset) {
    if ((e == 1 || e == 2 || e == 4 || e == 8) && D % (4 * e) == 0 && resident_bytes(ts, D / e, e) <= budget) chosen = e;
  } else if (chosen) {
    const int sms = num_sms();
    while (chosen < kMaxCluster && T * chosen < 2 * sms && D % (8 * chosen) == 0 && D / (2 * chosen) >= 256 &&
           resident_bytes(ts, D / (2 * chosen), 2 * chosen) <= budget)
      chosen *= 2;
  }
  if (!chosen) return false;
  pl->ts = ts;
  pl->cluster = chosen;
  pl->Ds = D / chosen;
  pl->smem = resident_bytes(ts, pl->Ds, chosen);
  return true;
}

template <int TS>
static int resident_launch(const ResidentPlan& pl, const float* emb, const float* dist, int T, int S, int D,
                           const scl_ms_params& p, float* per_tuple, float* demb, uint32_t* kept, float* loss,
                           unsigned int* counter, cudaStream_t stream) {
  auto kern = wms_tuple_kernel<TS>;
  static SmemAttrCache configured;            // per device; set only when it has to grow
  int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), pl.smem, &configured);
  if (rc_attr) return rc_attr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(T) * pl.cluster);
  cfg.blockDim = dim3(kRThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SCL_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, emb, dist, T, S, D, pl.Ds, p, per_tuple, demb, kept, loss, counter));
  return SCL_OK;
}

// SCL_ERR_UNSUPPORTED: the shape does not fit the resident kernel (caller falls through to the chunked one).
int wms_resident_launch(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params& p, float* loss,
                        float* per_tuple, float* demb, uint32_t* kept, unsigned int* counter, cudaStream_t stream) {
  if (knob_or(KNOB_WMS_CHUNKED, 0) != 0) return SCL_ERR_UNSUPPORTED;
  ResidentPlan pl;
  if (!resident_plan(T, S, D, &pl)) return SCL_ERR_UNSUPPORTED;
  switch (pl.ts) {
    case 5: return resident_launch<5>(pl, emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
    case 6: return resident_launch<6>(pl, emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
    default: return resident_launch<7>(pl, emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
  }
}

}  // namespace scl
