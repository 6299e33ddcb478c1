// wms_tuple_pipe.cu -- W1 tuple mode, warp-specialised throughput path for S <= 25: fused forward + analytic backward of
// the weighted multi-similarity loss (/root/reference/model/losses.py:5-60, call train/train.py:852, autodiff
// train.py:874-878), one persistent CTA per SM, THREE roles:
//
//   producer warp     streams [S x 256]-column tiles through a 7-stage shared-memory ring (cp.async.bulk per row);
//   8 streaming warps A: Gram E E^T with large register tiles ("stars": a lane owns two symmetric 5x5 tiles that share
//                        their x rows -- 15 row loads feed 200 FMAs per column quad), partials published per warp;
//                     C: dE = M E, 13 rows x 8 columns per thread (21 shared-memory wavefronts per 52 FFMA2);
//   3 weight warps    B: reduce the partial Grams, GPS masks, l2-norm, mining, log-sum-exp weights, M (ms_row.cuh).
//
// The streaming warps never wait for phase B: while the weight warps work on tuple t, the streaming warps already run
// the Gram of tuple t+1 (order A(0) | A(1) C(0) | A(2) C(1) | ...), and every hand-off is an mbarrier, so the FP32
// pipes and the shared-memory pipe see an uninterrupted stream of work.  Compared with wms_tuple_stream.cu (two CTAs
// per SM, 96 registers, small tiles) this trades occupancy for register tiles: 29 % fewer shared-memory wavefronts
// per tuple and no phase-B bubble.  Numerics, summation order and outputs are identical in structure (deterministic).
#include <atomic>
#include <cstdlib>

#include "ms_row.cuh"
#include "tc_common.cuh"
#include "tuple_common.cuh"

namespace scl {

using namespace tc;

namespace pipe {
constexpr int TS = 5, G = 5, SG = 25;                    // 5 x 5 grid of 5 x 5 tiles, S <= 25
constexpr int TILES = G * (G + 1) / 2, NP = TILES * TS * TS;
constexpr int NSTREAM = 8, NWEIGHT = 3;
constexpr int PRODUCER = NSTREAM;                        // warp index
constexpr int THREADS = (NSTREAM + 1 + NWEIGHT) * 32;    // 384
constexpr int CH = 256, PITCH = CH + 4;                  // columns per ring stage, row pitch in floats
constexpr int STAGES = 7;
constexpr int STAGE = SG * PITCH;
constexpr int HR = 13, HRP = 16, MTS = 2 * HRP;          // backward row halves, Mt row stride
constexpr int SQ = 628;                                  // al4(25 * 25)
constexpr int RB = 3;                                    // anchor rows per weight warp per batch
constexpr size_t FLOATS = size_t(STAGES) * STAGE + size_t(NSTREAM) * NP + 376 + 2 * SQ /*Gw, Sr*/ + 2 * SG * MTS /*Mt x2*/ +
                          2 * SQ /*dist x2*/ + 32 /*rowloss*/ + 32 /*invn*/;
constexpr int NBAR = 2 * STAGES + 2 + 2 + 1 + 1 + 2;
constexpr size_t BYTES = FLOATS * sizeof(float) + NBAR * sizeof(uint64_t);

__constant__ unsigned char c_star_c[8] = {0, 0, 1, 1, 2, 3, 4, 2};
__constant__ unsigned char c_star_y0[8] = {0, 2, 1, 3, 2, 3, 0, 4};
__constant__ unsigned char c_star_y1[8] = {1, 3, 2, 4, 3, 4, 4, 4};   // star 7 carries the single tile (2,4)

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void stg_hint(float4* ptr, const float4& v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void weight_sync() { asm volatile("bar.sync 2, %0;" ::"n"(NWEIGHT * 32) : "memory"); }

// ring order of the backward's re-read: chunk pairs descending (the most recently streamed columns first), ascending
// inside a pair
__device__ __forceinline__ int bwd_chunk(int kk, int nchunks) {
  const int npairs = (nchunks + 1) / 2, top = npairs - 1, first = nchunks - 2 * top;
  if (kk < first) return 2 * top + kk;
  const int r = kk - first;
  return 2 * (top - 1 - r / 2) + (r & 1);
}

// R rows x 8 columns of dE = M E: o[r] += M[r0 + r][j] * E[j][two column quads] over the S rows j
template <int R>
__device__ __forceinline__ void bwd_tile8(const float* __restrict__ ecol0, const float* __restrict__ ecol1,
                                          const float* __restrict__ mcol, int S, float2 (&o)[R][4]) {
#pragma unroll
  for (int r = 0; r < R; ++r) o[r][0] = o[r][1] = o[r][2] = o[r][3] = make_float2(0.0f, 0.0f);
#pragma unroll 5
  for (int j = 0; j < S; ++j) {
    const float4 e0 = *reinterpret_cast<const float4*>(ecol0 + j * PITCH);
    const float4 e1 = *reinterpret_cast<const float4*>(ecol1 + j * PITCH);
    const float* mrow = mcol + j * MTS;
    float mv[R];
#pragma unroll
    for (int r4 = 0; r4 < R / 4; ++r4) {
      const float4 m = *reinterpret_cast<const float4*>(mrow + 4 * r4);
      mv[4 * r4] = m.x; mv[4 * r4 + 1] = m.y; mv[4 * r4 + 2] = m.z; mv[4 * r4 + 3] = m.w;
    }
#pragma unroll
    for (int r = (R / 4) * 4; r < R; ++r) mv[r] = mrow[r];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ffma2(o[r][0], make_float2(mv[r], mv[r]), make_float2(e0.x, e0.y));
      ffma2(o[r][1], make_float2(mv[r], mv[r]), make_float2(e0.z, e0.w));
      ffma2(o[r][2], make_float2(mv[r], mv[r]), make_float2(e1.x, e1.y));
      ffma2(o[r][3], make_float2(mv[r], mv[r]), make_float2(e1.z, e1.w));
    }
  }
}

template <int R>
__device__ __forceinline__ void bwd_store8(float* __restrict__ dcol, int row0, int S, int D, bool v0, bool v1,
                                           const float2 (&o)[R][4], uint64_t pol) {
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (row0 + r < S) {
      float* dst = dcol + size_t(row0 + r) * D;
      if (v0) stg_hint(reinterpret_cast<float4*>(dst), make_float4(o[r][0].x, o[r][0].y, o[r][1].x, o[r][1].y), pol);
      if (v1) stg_hint(reinterpret_cast<float4*>(dst + 128), make_float4(o[r][2].x, o[r][2].y, o[r][3].x, o[r][3].y), pol);
    }
}

__global__ void __launch_bounds__(THREADS, 1) wms_pipe_kernel(
    const float* __restrict__ emb, const float* __restrict__ dist, int T, int S, int D, scl_ms_params p,
    float* __restrict__ per_tuple, float* __restrict__ demb, uint32_t* __restrict__ kept, float* __restrict__ loss_out,
    unsigned int* __restrict__ done_counter) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float* ring = smem;                                   // [STAGES][SG][PITCH]
  float* red = ring + size_t(STAGES) * STAGE;           // [NSTREAM][NP] partial Grams, tile order
  float* Pg = red + NSTREAM * NP;                       // [NP] reduced Gram (376)
  float* Gw = Pg + 376;                                 // dL/ds [SG][SG]
  float* Sr = Gw + SQ;                                  // raw cosine similarities [SG][SG]
  float* Mt = Sr + SQ;                                  // [2][SG][MTS]: Mt[j][h*HRP + r] = M[h*HR + r][j]
  float* dsm = Mt + 2 * SG * MTS;                       // [2][SQ] GPS distances
  float* rowloss = dsm + 2 * SQ;
  float* invn = rowloss + 32;
  uint64_t* full = reinterpret_cast<uint64_t*>(invn + 32);
  uint64_t* empty = full + STAGES;
  uint64_t* dfull = empty + STAGES;                     // [2]
  uint64_t* dempty = dfull + 2;                         // [2]
  uint64_t* gram_ready = dempty + 2;
  uint64_t* red_free = gram_ready + 1;
  uint64_t* m_ready = red_free + 1;                     // [2]

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NSTREAM);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&dfull[b], 32);                         // one cp.async-completion arrival per producer lane
      mbar_init(&dempty[b], NWEIGHT);
      mbar_init(&m_ready[b], 1);
    }
    mbar_init(gram_ready, NSTREAM);
    mbar_init(red_free, 1);
    fence_barrier_init();
  }
  // padding rows (>= S) of every stage stay zero for the whole kernel: the copies only write rows < S
  for (int s = 0; s < STAGES; ++s)
    for (int i = S * PITCH + tid; i < SG * PITCH; i += THREADS) ring[size_t(s) * STAGE + i] = 0.0f;
  __syncthreads();

  const int nchunks = (D + CH - 1) / CH;
  const bool need_bwd = demb != nullptr;
  const int n_mine = (T - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  // ============================ producer warp ============================
  if (warp == PRODUCER) {
    const uint64_t pol_keep = policy_evict_last(), pol_drop = policy_evict_first();
    const int kdist = nchunks > 2 ? 2 : nchunks - 1;    // distance block: after the first chunks are in flight
    uint32_t pos = 0;
    auto emit = [&](int t, int ch, uint64_t pol) {
      const int stage = pos % STAGES;
      mbar_wait(&empty[stage], ((pos / STAGES) & 1) ^ 1);
      const int c0 = ch * CH;
      const uint32_t bytes = uint32_t(min(CH, D - c0)) * 4u;
      if (lane == 0) mbar_arrive_expect_tx(&full[stage], bytes * uint32_t(S));
      __syncwarp();
      if (lane < S)
        bulk_load(ring + size_t(stage) * STAGE + lane * PITCH, emb + (size_t(t) * S + lane) * D + c0, bytes, &full[stage], pol);
      ++pos;
    };
    for (int s = 0; s <= n_mine; ++s) {
      if (s < n_mine) {                                 // phase A of tuple s
        const int t = blockIdx.x + s * gridDim.x;
        for (int k = 0; k < nchunks; ++k) {
          if (k == kdist) {
            const int b = s & 1;
            mbar_wait(&dempty[b], ((s >> 1) & 1) ^ 1);
            const float* dsrc = dist + size_t(t) * S * S;   // 4-byte cp.async: S*S*4 is not a multiple of 16 for odd S
            for (int i = lane; i < S * S; i += 32)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dsm + b * SQ + i)), "l"(dsrc + i) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&dfull[b])) : "memory");
          }
          emit(t, k, need_bwd ? pol_keep : pol_drop);
        }
      }
      if (s >= 1 && need_bwd) {                         // phase C of tuple s-1
        const int t = blockIdx.x + (s - 1) * gridDim.x;
        for (int kk = 0; kk < nchunks; ++kk) emit(t, bwd_chunk(kk, nchunks), pol_drop);
      }
    }
    return;
  }

  // ============================ weight warps (phase B) ============================
  if (warp > PRODUCER) {
    const int bw = warp - PRODUCER - 1, btid = tid - (PRODUCER + 1) * 32;
    const bool jvalid = lane < S;
    const int jj = jvalid ? lane : 0;
    const float invS = 1.0f / float(S), invT = 1.0f / float(T);
    auto gram = [&](int i, int j) -> float {
      int a = i % G, r = i / G, b = j % G, q = j / G;
      if (a > b) { int x = a; a = b; b = x; x = r; r = q; q = x; }
      return Pg[(r * TS + q) * TILES + (a * G - (a * (a - 1)) / 2 + (b - a))];
    };
    for (int it = 0; it < n_mine; ++it) {
      const int t = blockIdx.x + it * gridDim.x;
      const int b = it & 1;
      const float* ds = dsm + b * SQ;
      float* Mt_b = Mt + b * SG * MTS;
      mbar_wait(gram_ready, it & 1);
      for (int k = btid; k < NP; k += NWEIGHT * 32) {   // fixed summation order: deterministic
        float v = red[k];
#pragma unroll
        for (int w = 1; w < NSTREAM; ++w) v += red[w * NP + k];
        Pg[k] = v;
      }
      weight_sync();
      if (btid == 0) mbar_arrive(red_free);
      // tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))  (losses.py:7); below the clamp it is a pure scale
      const float n2 = gram(jj, jj);
      const float invn_j = rsqrtf(fmaxf(n2, 1e-12f));
      const float nflag_j = n2 >= 1e-12f ? 1.0f : 0.0f;
      mbar_wait(&dfull[b], (it >> 1) & 1);
      // ---- pass 1: weights of the anchor rows bw, bw + 3, ... (RB rows interleaved), lane = column j ----
      for (int r0 = bw; r0 < S; r0 += NWEIGHT * RB) {
        float wpv[RB], wnv[RB], invn_i[RB], raw[RB], sv[RB];
        bool rvalid[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) {
          const int i = r0 + NWEIGHT * k;
          rvalid[k] = i < S;
          wpv[k] = wnv[k] = 0.0f;
          if (rvalid[k] && jvalid) {
            wms_masks(ds[i * S + lane], p.d_alpha, p.d_beta, p.wfunction, wpv[k], wnv[k]);   // losses.py:11-19
            if (i == lane) wpv[k] -= 1.0f;                                                   // losses.py:22
          }
          const int ii = rvalid[k] ? i : 0;
          invn_i[k] = __shfl_sync(0xffffffffu, invn_j, ii);
          raw[k] = (rvalid[k] && jvalid) ? gram(ii, jj) * invn_i[k] * invn_j : 0.0f;
          sv[k] = fmaxf(raw[k], 0.0f);                                                       // losses.py:26
        }
        MsRowStats st[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) {
          st[k].maxv = jvalid ? sv[k] * wnv[k] : -INFINITY;
          st[k].tmp = jvalid ? sv[k] * wpv[k] : -INFINITY;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int k = 0; k < RB; ++k) {
            st[k].maxv = fmaxf(st[k].maxv, __shfl_xor_sync(0xffffffffu, st[k].maxv, o));
            st[k].tmp = fmaxf(st[k].tmp, __shfl_xor_sync(0xffffffffu, st[k].tmp, o));
          }
#pragma unroll
        for (int k = 0; k < RB; ++k) st[k].minv = jvalid ? (sv[k] - st[k].tmp) * wpv[k] : INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int k = 0; k < RB; ++k) st[k].minv = fminf(st[k].minv, __shfl_xor_sync(0xffffffffu, st[k].minv, o));
        bool kp[RB], kn[RB];
        float ep[RB], en[RB], A[RB], B[RB];
#pragma unroll
        for (int k = 0; k < RB; ++k) {
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
          for (int k = 0; k < RB; ++k) {
            A[k] += __shfl_xor_sync(0xffffffffu, A[k], o);
            B[k] += __shfl_xor_sync(0xffffffffu, B[k], o);
          }
#pragma unroll
        for (int k = 0; k < RB; ++k) {
          const int i = r0 + NWEIGHT * k;
          if (rvalid[k]) {
            if (jvalid) {
              float gw = ms_elem_grad(wpv[k], wnv[k], kp[k], kn[k], ep[k], en[k], A[k], B[k], p) * invS;
              if (!(raw[k] >= 0.0f)) gw = 0.0f;                              // tf.maximum passes gradient when x >= 0
              Gw[i * SG + lane] = gw;
              Sr[i * SG + lane] = raw[k];
            }
            if (lane == 0) {
              rowloss[i] = ms_row_loss(A[k], B[k], p) * invS;
              invn[i] = invn_i[k];
            }
          }
          if (kept != nullptr) {
            const unsigned mp = __ballot_sync(0xffffffffu, kp[k]), mn = __ballot_sync(0xffffffffu, kn[k]);
            if (lane == 0 && rvalid[k]) {
              kept[(size_t(t) * S + i) * 2 + 0] = mp;
              kept[(size_t(t) * S + i) * 2 + 1] = mn;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[b]);
      weight_sync();
      // ---- pass 2: M = (1/T) diag(invn) (W - diag(c)) diag(invn), W = Gw + Gw^T, c_i = sum_j W_ij s_ij(raw) ----
      for (int i = bw; i < S; i += NWEIGHT) {
        const float wij = jvalid ? Gw[i * SG + lane] + Gw[lane * SG + i] : 0.0f;
        float c = wij * (jvalid ? Sr[i * SG + lane] : 0.0f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        c *= __shfl_sync(0xffffffffu, nflag_j, i);
        if (jvalid) {
          const float w = (i == lane) ? wij - c : wij;
          const int h = i / HR, r = i - h * HR;
          Mt_b[lane * MTS + h * HRP + r] = invn[i] * w * invn_j * invT;
        }
      }
      if (bw == 0) {
        float v = lane < S ? rowloss[lane] : 0.0f;
        v = warp_sum(v);
        if (lane == 0 && per_tuple != nullptr) per_tuple[t] = v;
        tup_deposit_loss(done_counter, t, v, lane);
      }
      weight_sync();
      if (btid == 0) mbar_arrive(&m_ready[b]);
    }
    if (bw == 0) tup_publish_losses(done_counter, n_mine, T, loss_out, lane);
    return;
  }

  // ============================ streaming warps (phases A and C) ============================
  const int star = lane & 7, cg = lane >> 3;               // 8 stars x 4 column groups per warp
  const int sc = c_star_c[star], sy0 = c_star_y0[star], sy1 = c_star_y1[star];
  int xo = sc * PITCH, y0o = sy0 * PITCH, y1o = sy1 * PITCH;
  asm volatile("" : "+r"(xo), "+r"(y0o), "+r"(y1o));       // opaque: row offsets stay in registers
  const uint64_t pol_drop = policy_evict_first();
  uint32_t pos = 0;
  for (int s = 0; s <= n_mine; ++s) {
    if (s < n_mine) {
      // ---------------- A. Gram of tuple s ----------------
      float2 acc0[TS][TS], acc1[TS][TS];
#pragma unroll
      for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int q = 0; q < TS; ++q) acc0[r][q] = acc1[r][q] = make_float2(0.0f, 0.0f);
#pragma unroll 1
      for (int ch = 0; ch < nchunks; ++ch, ++pos) {
        const int stage = pos % STAGES;
        const int nq = min(CH, D - ch * CH) >> 2;
        const float* Es = ring + size_t(stage) * STAGE;
        mbar_wait(&full[stage], (pos / STAGES) & 1);
#pragma unroll 1
        for (int c4 = warp * 4 + cg; c4 < nq; c4 += 4 * NSTREAM) {
          const float* xcol = Es + xo + 4 * c4;
          const float* y0col = Es + y0o + 4 * c4;
          const float* y1col = Es + y1o + 4 * c4;
          float4 x[TS];
#pragma unroll
          for (int r = 0; r < TS; ++r) x[r] = *reinterpret_cast<const float4*>(xcol + G * r * PITCH);
#pragma unroll
          for (int q = 0; q < TS; ++q) {
            const float4 y = *reinterpret_cast<const float4*>(y0col + G * q * PITCH);
#pragma unroll
            for (int r = 0; r < TS; ++r) ffma2(acc0[r][q], make_float2(x[r].x, x[r].y), make_float2(y.x, y.y));
#pragma unroll
            for (int r = 0; r < TS; ++r) ffma2(acc0[r][q], make_float2(x[r].z, x[r].w), make_float2(y.z, y.w));
          }
#pragma unroll
          for (int q = 0; q < TS; ++q) {
            const float4 y = *reinterpret_cast<const float4*>(y1col + G * q * PITCH);
#pragma unroll
            for (int r = 0; r < TS; ++r) ffma2(acc1[r][q], make_float2(x[r].x, x[r].y), make_float2(y.x, y.y));
#pragma unroll
            for (int r = 0; r < TS; ++r) ffma2(acc1[r][q], make_float2(x[r].z, x[r].w), make_float2(y.z, y.w));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      // fold even/odd columns and the warp's four column groups; lanes 0..7 publish the warp's partial stars in
      // canonical tile order (tile (a <= b): entry (r, q) = row(a + 5r) . row(b + 5q))
      if (s > 0) mbar_wait(red_free, (s - 1) & 1);          // the weight warps have summed the previous partials
      float* dst = red + warp * NP;
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        const int g = pr ? sy1 : sy0;
        const bool swap = sc > g;
        const int a = swap ? g : sc, b = swap ? sc : g;
        const int tix = a * G - (a * (a - 1)) / 2 + (b - a);
        const bool wr = lane < 8 && !(pr == 1 && star == 7);
#pragma unroll
        for (int r = 0; r < TS; ++r)
#pragma unroll
          for (int q = 0; q < TS; ++q) {
            float v = pr ? acc1[r][q].x + acc1[r][q].y : acc0[r][q].x + acc0[r][q].y;
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (wr) dst[(swap ? q * TS + r : r * TS + q) * TILES + tix] = v;
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(gram_ready);
    }
    if (s >= 1 && need_bwd) {
      // ---------------- C. backward of tuple s-1: demb = M E ----------------
      const int it = s - 1;
      const int t = blockIdx.x + it * gridDim.x;
      const float* Mt_b = Mt + (it & 1) * SG * MTS;
      float* dE_t = demb + size_t(t) * S * D;
      mbar_wait(&m_ready[it & 1], (it >> 1) & 1);
      // one item = (row half h, ring stage): a thread owns column quads `lane` and `lane + 32` of the stage; the two
      // items of a stage belong to two different warps, each releases the stage with half of the arrivals
#pragma unroll 1
      for (int kk = 0; kk < nchunks; ++kk) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (((2 * kk + h) & (NSTREAM - 1)) != warp) continue;
          const uint32_t ps = pos + kk;
          const int stage = ps % STAGES;
          const int ch = bwd_chunk(kk, nchunks);
          const int nqs = min(CH, D - ch * CH) >> 2;
          const float* ebase = ring + size_t(stage) * STAGE;
          mbar_wait(&full[stage], (ps / STAGES) & 1);
          const bool v0 = lane < nqs, v1 = lane + 32 < nqs;
          const float* e0 = ebase + 4 * (v0 ? lane : 0);
          const float* e1 = ebase + 4 * (v1 ? lane + 32 : 0);
          float* dcol = dE_t + size_t(ch) * CH + 4 * lane;
          if (h == 0) {
            float2 o[HR][4];
            bwd_tile8<HR>(e0, e1, Mt_b, S, o);
            bwd_store8<HR>(dcol, 0, S, D, v0, v1, o, pol_drop);
          } else {
            float2 o[SG - HR][4];
            bwd_tile8<SG - HR>(e0, e1, Mt_b + HRP, S, o);
            bwd_store8<SG - HR>(dcol, HR, S, D, v0, v1, o, pol_drop);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_n(&empty[stage], NSTREAM / 2);
        }
      }
      pos += nchunks;
    }
  }
}

}  // namespace pipe

// SCL_ERR_UNSUPPORTED: shapes outside S <= 25 or small batches stay with the other kernels.
int wms_pipe_launch(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params& p, float* loss,
                    float* per_tuple, float* demb, uint32_t* kept, unsigned int* counter, cudaStream_t stream) {
  if (S < 2 || S > pipe::SG || D < 4 || (D & 3)) return SCL_ERR_UNSUPPORTED;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_relaxed)) {
    SCL_CUDA_TRY(cudaFuncSetAttribute(pipe::wms_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pipe::BYTES)));
    configured.store(1, std::memory_order_relaxed);
  }
  int grid = num_sms();
  if (grid > T) grid = T;
  pipe::wms_pipe_kernel<<<grid, pipe::THREADS, pipe::BYTES, stream>>>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss,
                                                                     counter);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

}  // namespace scl
