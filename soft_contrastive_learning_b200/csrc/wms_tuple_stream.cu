// wms_tuple_stream.cu -- W1 tuple mode, throughput path: fused forward + analytic backward of the weighted
// multi-similarity loss for large batches of tuples, one tuple per CTA at a time, descriptors STREAMED through a
// shared-memory ring by a dedicated producer warp.
//
// Replaces wms_loss (/root/reference/model/losses.py:5-60, call train/train.py:852) and the TF autodiff of it
// (train.py:874-878) for T independent tuples of S <= 32 descriptors of any dimension D (multiple of 4).
//
// Persistent CTAs (two per SM); CTA b handles tuples b, b + grid, ...  Per tuple the eight consumer warps run
//   A. Gram: the S x D tile arrives in 256-column chunks (cp.async.bulk per row, mbarrier per ring stage); the
//      symmetric 5x5 grid of TSxTS register tiles accumulates E E^T with packed FFMA2 (even/odd columns in one
//      register pair); a two-round cross-warp reduction leaves the S x S Gram in shared memory;
//   B. weights: GPS soft masks (losses.py:11-22) from the prefetched distance block, l2-normalisation, mining
//      thresholds, log-sum-exp weights (ms_row.cuh), four anchor rows interleaved per warp; the l2-norm Jacobian
//      and the 1/T batch mean are folded into one S x S matrix M;
//   C. backward: the tile is streamed a second time (last chunk first, so the re-read finds it in L2) and
//      d loss / d emb = M E leaves as 512-byte-per-row coalesced streaming stores; M enters FFMA2 as a broadcast
//      scalar operand.
// The producer warp runs ahead of the consumers across phases and tuples (the first chunks of tuple t+1 are in
// flight while tuple t finishes), so neither the scalar phase B nor the load latency leaves the FP32 pipes idle as
// long as the other CTA of the SM has work.  No cluster, no cross-CTA exchange, nothing computed redundantly.
// HBM traffic is the algorithmic minimum as long as the second read hits L2: emb read once, demb written once.
#include <atomic>
#include <cstdlib>

#include "ms_row.cuh"
#include "tc_common.cuh"
#include "tuple_common.cuh"

namespace scl {

using namespace tc;

constexpr int kSGrid = 5;                                  // 5 x 5 grid of TS x TS tiles
constexpr int kSTiles = kSGrid * (kSGrid + 1) / 2;         // 15 symmetric tiles
constexpr int kSStages = 4;
constexpr int kSRed = 4;                                   // cross-warp reduction buffers (two rounds)

// NW consumer warps (+ one producer warp), CH columns per ring stage.  Two configurations are instantiated:
//   <8, 256>  two CTAs per SM  -- two tuples in flight per SM: 2 x 148 x 410 KB of live descriptors do not fit in L2
//                                 next to the gradient stream, so ~60 % of the backward's re-read comes from HBM;
//   <16, 512> one CTA per SM   -- one tuple in flight per SM (60 MB live): the re-read hits L2; 17 warps are
//                                 allocated as 20, which caps the kernel at 96 registers per thread;
//   <15, 480> one CTA per SM   -- the same with 16 warps in total: 128 registers per thread;
//   <7, 224>  two CTAs per SM  -- 8 warps in total per CTA: 128 registers per thread instead of 96.
constexpr int kMpStride = 36;                              // M for the tensor-core backward: [32][36] floats (bank = 4*row + col)
template <int TS, int CH, bool kMma = false>
struct SSmem {
  // floats; rows 16 bytes apart in bank space (FFMA2 paths), 48 bytes for the tensor-core backward (its scalar
  // fragment loads, bank = 12*(lane%4) + lane/4, and the Gram's vector loads are both conflict-free)
  static constexpr int kSPitch = CH + (kMma ? 12 : 4);
  static constexpr int SG = kSGrid * TS;
  static constexpr int NP = kSTiles * TS * TS;
  static constexpr int HR = (SG + 1) / 2;
  static constexpr int HRP = (HR + 3) / 4 * 4;
  static constexpr int MT = kMma ? 32 * kMpStride : SG * 2 * HRP;
  static constexpr int STAGE = SG * kSPitch;                                       // floats per ring stage
  static constexpr int GW = int(al4(SG * SG));
  static constexpr int WORK = (kSRed * NP > GW + MT) ? kSRed * NP : GW + MT;       // reduction, then Gw | Mt
  static constexpr int DIST = int(al4(SG * SG));
  static constexpr size_t floats = size_t(kSStages) * STAGE + al4(NP) + al4(WORK) + DIST + 32;
  static constexpr size_t bytes = floats * sizeof(float) + (2 * kSStages + 2) * sizeof(uint64_t);
};

__constant__ unsigned char c_stile_a[kSTiles] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4};
__constant__ unsigned char c_stile_b[kSTiles] = {0, 1, 2, 3, 4, 1, 2, 3, 4, 2, 3, 4, 3, 4, 4};

__device__ __forceinline__ void s_ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void s_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// L2 eviction policies: the first read of a tuple should survive until the backward re-reads it (evict_last); the
// re-read and the gradient stores are dead on arrival (evict_first) and must not push the live tuples out
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
__device__ __forceinline__ void stg_hint2_if(bool on, float* ptr, float x, float y, uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t@p st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;\n\t}" ::"l"(ptr),
      "f"(x), "f"(y), "l"(policy), "r"(int(on))
      : "memory");
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int N>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// R rows x 4 columns of dE = M E for one column quad: o[r] += M[r0 + r][j] * E[j][c4*4 .. +3] over the S rows j.
// An LDS.128 costs four shared-memory wavefronts whether or not its lanes agree on the address, so the broadcast M
// operands are fetched with exactly ceil(R/4) vector loads (+ scalar loads for the remainder), never for padding.
template <int R, int PITCH, int MSTRIDE>
__device__ __forceinline__ void bwd_tile(const float* __restrict__ ecol, const float* __restrict__ mcol, int S,
                                         float2 (&o)[R][2]) {
#pragma unroll
  for (int r = 0; r < R; ++r) o[r][0] = o[r][1] = make_float2(0.0f, 0.0f);
#pragma unroll 5
  for (int j = 0; j < S; ++j) {
    const float4 e = *reinterpret_cast<const float4*>(ecol + j * PITCH);
    const float* mrow = mcol + j * MSTRIDE;
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
      s_ffma2(o[r][0], make_float2(mv[r], mv[r]), make_float2(e.x, e.y));
      s_ffma2(o[r][1], make_float2(mv[r], mv[r]), make_float2(e.z, e.w));
    }
  }
}

template <int TS, int NW, int CH, bool kPackedGram = true, bool kMmaBwd = false>
__global__ void __launch_bounds__(NW * 32 + 32, ((TS == 5 && NW <= 8) ? 2 : 1)) wms_stream_kernel(
    const float* __restrict__ emb, const float* __restrict__ dist, int T, int S, int D, scl_ms_params p,
    float* __restrict__ per_tuple, float* __restrict__ demb, uint32_t* __restrict__ kept, float* __restrict__ loss_out,
    unsigned int* __restrict__ done_counter) {
  using L = SSmem<TS, CH, kMmaBwd>;
  constexpr int SG = L::SG, NP = L::NP, HR = L::HR, HRP = L::HRP, STAGE = L::STAGE;
  constexpr int kSWarps = NW, kSConsumers = NW * 32, kSThreads = NW * 32 + 32, kSChunk = CH, kSPitch = L::kSPitch;
  constexpr int RPW = (32 + NW - 1) / NW;            // anchor rows per warp in phase B (S <= 32)
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float* ring = smem;                                   // [kSStages][SG][kSPitch]
  float* Pg = ring + size_t(kSStages) * STAGE;          // reduced Gram in tile order [TS*TS][15]
  float* work = Pg + al4(NP);                           // reduction buffers, later Gw [SG][SG] | Mt [SG][2*HRP]
  float* dsm = work + al4(L::WORK);                     // this tuple's GPS distances [S][S]
  float* rowloss = dsm + L::DIST;
  uint64_t* full = reinterpret_cast<uint64_t*>(rowloss + 32);
  uint64_t* empty = full + kSStages;
  uint64_t* dfull = empty + kSStages;
  uint64_t* dempty = dfull + 1;

  if (tid == 0) {
    for (int s = 0; s < kSStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kSWarps);
    }
    mbar_init(dfull, 32);                                // one cp.async-completion arrival per producer lane
    mbar_init(dempty, kSWarps);
    fence_barrier_init();
  }
  // padding rows (>= S) of every stage stay zero for the whole kernel: the copies only write rows < S
  for (int s = 0; s < kSStages; ++s)
    for (int i = S * kSPitch + tid; i < SG * kSPitch; i += kSThreads) ring[size_t(s) * STAGE + i] = 0.0f;
  __syncthreads();

  const int nchunks = (D + kSChunk - 1) / kSChunk;
  const int npairs = (nchunks + 1) / 2;
  const bool need_bwd = demb != nullptr;

  // ============================ producer warp ============================
  // Ring positions run across phases and tuples: per tuple nchunks positions for phase A (chunks ascending) and, with
  // a backward, nchunks more for phase C (pairs descending, ascending inside a pair).  (Folding this role into a
  // consumer warp was measured 25 % slower: that warp blocks on the ring's empty barriers and becomes the straggler.)
  if (warp == kSWarps) {
    const int nseq = nchunks + (need_bwd ? nchunks : 0);
    const int kdist = nseq > 2 ? 2 : nseq - 1;          // distance block: after the first chunks are in flight
    const uint64_t pol_keep = policy_evict_last(), pol_drop = policy_evict_first();
    uint32_t pos = 0, iter = 0;
    for (int t = blockIdx.x; t < T; t += gridDim.x, ++iter) {
      for (int k = 0; k < nseq; ++k, ++pos) {
        if (k == kdist) {
          // GPS distances of this tuple (single buffer: free once phase B of the previous tuple has read it);
          // 4-byte cp.async because S*S*4 is not a multiple of 16 for odd S
          mbar_wait(dempty, (iter & 1) ^ 1);
          const float* dsrc = dist + size_t(t) * S * S;
          for (int i = lane; i < S * S; i += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dsm + i)), "l"(dsrc + i) : "memory");
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(dfull)) : "memory");
        }
        int ch;
        if (k < nchunks) {
          ch = k;
        } else {
          const int kk = k - nchunks;                   // 0 .. nchunks-1
          const int top = npairs - 1;                   // the (possibly half-filled) last pair comes first
          const int first = nchunks - 2 * top;          // chunks in the last pair: 1 or 2
          if (kk < first) ch = 2 * top + kk;
          else { const int r = kk - first; ch = 2 * (top - 1 - r / 2) + (r & 1); }
        }
        const int stage = pos % kSStages;
        mbar_wait(&empty[stage], ((pos / kSStages) & 1) ^ 1);
        const int c0 = ch * kSChunk;
        const uint32_t bytes = uint32_t(min(kSChunk, D - c0)) * 4u;
        if (lane == 0) mbar_arrive_expect_tx(&full[stage], bytes * uint32_t(S));
        __syncwarp();
        if (lane < S)
          s_bulk_load(ring + size_t(stage) * STAGE + lane * kSPitch, emb + (size_t(t) * S + lane) * D + c0, bytes, &full[stage],
                      (k < nchunks && need_bwd) ? pol_keep : pol_drop);
      }
    }
    return;
  }

  // ============================ consumer warps ============================
  const int tl = lane % kSTiles;
  const int dg = lane / kSTiles;                   // 0,1 (lanes 30,31 -> 2: idle in the Gram)
  const int ta = c_stile_a[tl], tb = c_stile_b[tl];
  // the warp's two column groups sit kSWarps quads apart (8 quads = 128 bytes: the same banks, so the 5 rows of one
  // group never collide with a row of the other inside a quarter-warp phase)
  const int dgid = warp + dg * kSWarps;             // 2 * kSWarps column groups per CTA
  const bool jvalid = lane < S;
  const int jj = jvalid ? lane : 0;
  const float invS = 1.0f / float(S);
  const float invT = 1.0f / float(T);
  float* Gw = work;                                // dL/ds_ij [SG][SG]
  float* Mt = work + L::GW;                        // [SG][2*HRP], 16-byte aligned
  uint32_t pos = 0;
  uint32_t iter = 0;
  const uint64_t pol_drop = policy_evict_first();

  for (int t = blockIdx.x; t < T; t += gridDim.x, ++iter) {
    // ---------------- A. Gram ----------------
    // packed: even/odd-column partial sums in one register pair (FFMA2); scalar: one accumulator per pair (FFMA), half
    // the accumulator registers, which lets the compiler keep the operand loads further ahead of their use
    float2 acc[TS][TS];
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
      for (int q = 0; q < TS; ++q) acc[r][q] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int ch = 0; ch < nchunks; ++ch, ++pos) {
      const int stage = pos % kSStages;
      const int nq = min(kSChunk, D - ch * kSChunk) >> 2;
      const float* Es = ring + size_t(stage) * STAGE;
      mbar_wait(&full[stage], (pos / kSStages) & 1);
      if (dg < 2) {
#pragma unroll 1
        for (int c4 = dgid; c4 < nq; c4 += 2 * kSWarps) {
          // y rows stay in registers for the iteration, x rows are double-buffered one ahead of their use
          const float* xcol = Es + ta * kSPitch + 4 * c4;
          const float* ycol = Es + tb * kSPitch + 4 * c4;
          float4 y[TS];
#pragma unroll
          for (int r = 0; r < TS; ++r) y[r] = *reinterpret_cast<const float4*>(ycol + kSGrid * r * kSPitch);
          float4 xn = *reinterpret_cast<const float4*>(xcol);
#pragma unroll
          for (int r = 0; r < TS; ++r) {
            const float4 x = xn;
            if (r + 1 < TS) xn = *reinterpret_cast<const float4*>(xcol + kSGrid * (r + 1) * kSPitch);
            if (kPackedGram) {
#pragma unroll
              for (int q = 0; q < TS; ++q) s_ffma2(acc[r][q], make_float2(x.x, x.y), make_float2(y[q].x, y[q].y));
#pragma unroll
              for (int q = 0; q < TS; ++q) s_ffma2(acc[r][q], make_float2(x.z, x.w), make_float2(y[q].z, y[q].w));
            } else {
#pragma unroll
              for (int q = 0; q < TS; ++q) acc[r][q].x = fmaf(x.x, y[q].x, acc[r][q].x);
#pragma unroll
              for (int q = 0; q < TS; ++q) acc[r][q].x = fmaf(x.y, y[q].y, acc[r][q].x);
#pragma unroll
              for (int q = 0; q < TS; ++q) acc[r][q].x = fmaf(x.z, y[q].z, acc[r][q].x);
#pragma unroll
              for (int q = 0; q < TS; ++q) acc[r][q].x = fmaf(x.w, y[q].w, acc[r][q].x);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    // fold even/odd columns, the two column groups of a warp, then the warps (two rounds over kSRed buffers)
    consumer_sync<kSConsumers>();                               // nobody still reads Mt of the previous tuple (tiny D)
    {
      float g[TS][TS];
#pragma unroll
      for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int q = 0; q < TS; ++q) {
          const float v = acc[r][q].x + acc[r][q].y;
          g[r][q] = v + __shfl_down_sync(0xffffffffu, v, kSTiles);
        }
      // groups of kSRed warps take turns, highest warps first: write, then add-and-write, ... (fixed order: deterministic)
      const int wrev = kSWarps - 1 - warp;
#pragma unroll
      for (int grp = 0; grp < (kSWarps + kSRed - 1) / kSRed; ++grp) {
        if (wrev / kSRed == grp && lane < kSTiles) {
          float* dst = work + (wrev % kSRed) * NP;
#pragma unroll
          for (int r = 0; r < TS; ++r)
#pragma unroll
            for (int q = 0; q < TS; ++q) {
              const int a = (r * TS + q) * kSTiles + tl;
              dst[a] = (grp == 0) ? g[r][q] : dst[a] + g[r][q];
            }
        }
        consumer_sync<kSConsumers>();
      }
      for (int k = tid; k < NP; k += kSConsumers) Pg[k] = (work[k] + work[NP + k]) + (work[2 * NP + k] + work[3 * NP + k]);
      consumer_sync<kSConsumers>();
    }

    // ---------------- B. weights: warp w owns anchor rows w, w+8, w+16, w+24 (interleaved), lane = column j ----------------
    auto gram = [&](int i, int j) -> float {
      int a = i % kSGrid, r = i / kSGrid, b = j % kSGrid, q = j / kSGrid;
      if (a > b) { int x = a; a = b; b = x; x = r; r = q; q = x; }
      return Pg[(r * TS + q) * kSTiles + (a * kSGrid - (a * (a - 1)) / 2 + (b - a))];
    };
    mbar_wait(dfull, iter & 1);
    float wpv[RPW], wnv[RPW];
    bool rvalid[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      const int i = warp + kSWarps * k;
      rvalid[k] = i < S;
      wpv[k] = 0.0f;
      wnv[k] = 0.0f;
      if (rvalid[k] && jvalid) {
        wms_masks(dsm[i * S + lane], p.d_alpha, p.d_beta, p.wfunction, wpv[k], wnv[k]);   // losses.py:11-19
        if (i == lane) wpv[k] -= 1.0f;                                                     // losses.py:22
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(dempty);
    const float n2 = gram(jj, jj);
    // tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))  (losses.py:7); below the clamp it is a pure scale
    const float invn_j = rsqrtf(fmaxf(n2, 1e-12f));
    const float nflag_j = n2 >= 1e-12f ? 1.0f : 0.0f;
    float invn_i[RPW], raw[RPW], sv[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      const int ii = rvalid[k] ? warp + kSWarps * k : 0;
      invn_i[k] = __shfl_sync(0xffffffffu, invn_j, ii);
      raw[k] = (rvalid[k] && jvalid) ? gram(ii, jj) * invn_i[k] * invn_j : 0.0f;
      sv[k] = fmaxf(raw[k], 0.0f);                                           // losses.py:26
    }
    MsRowStats st[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      st[k].maxv = jvalid ? sv[k] * wnv[k] : -INFINITY;
      st[k].tmp = jvalid ? sv[k] * wpv[k] : -INFINITY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        st[k].maxv = fmaxf(st[k].maxv, __shfl_xor_sync(0xffffffffu, st[k].maxv, o));
        st[k].tmp = fmaxf(st[k].tmp, __shfl_xor_sync(0xffffffffu, st[k].tmp, o));
      }
#pragma unroll
    for (int k = 0; k < RPW; ++k) st[k].minv = jvalid ? (sv[k] - st[k].tmp) * wpv[k] : INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < RPW; ++k) st[k].minv = fminf(st[k].minv, __shfl_xor_sync(0xffffffffu, st[k].minv, o));
    bool kp[RPW], kn[RPW];
    float ep[RPW], en[RPW], A[RPW], B[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
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
      for (int k = 0; k < RPW; ++k) {
        A[k] += __shfl_xor_sync(0xffffffffu, A[k], o);
        B[k] += __shfl_xor_sync(0xffffffffu, B[k], o);
      }
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      const int i = warp + kSWarps * k;
      if (rvalid[k]) {
        if (jvalid) {
          float gw = ms_elem_grad(wpv[k], wnv[k], kp[k], kn[k], ep[k], en[k], A[k], B[k], p) * invS;
          if (!(raw[k] >= 0.0f)) gw = 0.0f;                                  // tf.maximum passes gradient when x >= 0
          Gw[i * SG + lane] = gw;
        }
        if (lane == 0) rowloss[i] = ms_row_loss(A[k], B[k], p) * invS;
      }
      if (kept != nullptr) {
        const unsigned mp = __ballot_sync(0xffffffffu, kp[k]), mn = __ballot_sync(0xffffffffu, kn[k]);
        if (lane == 0 && rvalid[k]) {
          kept[(size_t(t) * S + i) * 2 + 0] = mp;
          kept[(size_t(t) * S + i) * 2 + 1] = mn;
        }
      }
    }
    consumer_sync<kSConsumers>();
    // M = (1/T) diag(invn) (W - diag(c)) diag(invn), W = Gw + Gw^T, c_i = sum_j W_ij s_ij(raw)  (projection of l2norm),
    // stored transposed and split in two row halves for the backward: Mt[j][h*HRP + r] = M[h*HR + r][j]
    {
      float wij[RPW], cpart[RPW];
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int i = warp + kSWarps * k;
        wij[k] = (rvalid[k] && jvalid) ? Gw[i * SG + lane] + Gw[lane * SG + i] : 0.0f;
        cpart[k] = wij[k] * raw[k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < RPW; ++k) cpart[k] += __shfl_xor_sync(0xffffffffu, cpart[k], o);
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int i = warp + kSWarps * k;
        const float c = cpart[k] * __shfl_sync(0xffffffffu, nflag_j, rvalid[k] ? i : 0);
        const float w = (i == lane) ? wij[k] - c : wij[k];
        if constexpr (kMmaBwd) {
          // plain [32][kMpStride] layout, zero outside S x S (the padded tensor-core tiles multiply it)
          if (i < 32) Mt[i * kMpStride + lane] = (rvalid[k] && jvalid) ? invn_i[k] * w * invn_j * invT : 0.0f;
        } else if (rvalid[k] && jvalid) {
          const int h = i / HR, r = i - h * HR;
          Mt[lane * (2 * HRP) + h * HRP + r] = invn_i[k] * w * invn_j * invT;
        }
      }
    }
    // ---------------- loss ----------------
    if (warp == 0) {
      float v = lane < S ? rowloss[lane] : 0.0f;
      v = warp_sum(v);
      if (lane == 0 && per_tuple != nullptr) per_tuple[t] = v;
      tup_deposit_loss(done_counter, t, v, lane);            // published once per CTA after the last tuple
    }
    consumer_sync<kSConsumers>();

    // ---------------- C. backward: demb = M * E, chunk pairs in descending order ----------------
    if constexpr (kMmaBwd) { if (need_bwd) {
      // Tensor-core backward: dE[32 x cols] = M[32 x 8*KT] E[8*KT x cols] on mma.sync.m16n8k8 tf32 with the 3-term
      // hi/lo split (hi*hi + lo*hi + hi*lo; the dropped lo*lo is 2^-22 of a product: fp32-grade).  It moves the
      // backward's 2.56 M FMAs per tuple off the FP32 pipes (they keep the Gram of the SM's other CTA) and most of its
      // operand traffic off the shared-memory pipe (7 loads per 8 columns instead of 25 vector loads per 4).  M's
      // fragments (both 16-row halves, hi and lo) live in registers for the tuple; a warp owns 32-column groups of a
      // stage: E is loaded and split once per 8 columns and feeds four independent accumulator chains (main = hi*hi
      // and correction = lo*hi + hi*lo, per row half).  S == 25: 24 rows of E go through the tensor core, row 24 is
      // a rank-1 FFMA update folded into the accumulator initialisation.
      // Measured alternatives (B200, T = 4096): the transposed product (E^T as the 16-row operand, 27 instead of 36
      // MMAs per 16 columns, 4-byte stores) 1.03 ms; interleaved even/odd column tiles with 16-byte stores 0.98 ms;
      // this form 0.96 ms; the FFMA2 backward 1.07 ms.
      constexpr int KT = (SG <= 25) ? 3 : 4;
      const int g = lane >> 2, tg = lane & 3;
      uint32_t ahi[2][KT][4], alo[2][KT][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v = Mt[(16 * mt + g + 8 * (e & 1)) * kMpStride + 8 * kt + tg + 4 * (e >> 1)];
            ahi[mt][kt][e] = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
            alo[mt][kt][e] = __float_as_uint(v - __uint_as_float(ahi[mt][kt][e]));
          }
      float m24[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) m24[e] = (KT == 3) ? Mt[(g + 8 * e) * kMpStride + 24] : 0.0f;
      int roff[KT][2];                                  // E rows of the B fragments (clamped into the stage: M is 0 there)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int h = 0; h < 2; ++h) roff[kt][h] = min(8 * kt + 4 * h + tg, SG - 1) * kSPitch + g;
      float* dbase = demb + size_t(t) * S * D + size_t(g) * D + 2 * tg;
      const size_t D8 = size_t(8) * D;
      bool stv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) stv[e] = g + 8 * e < S;
      auto load_b = [&](const float* Es, int c0, float (&b)[KT][2], float2& e24) {
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          b[kt][0] = Es[roff[kt][0] + c0];
          b[kt][1] = Es[roff[kt][1] + c0];
        }
        if (KT == 3) e24 = *reinterpret_cast<const float2*>(Es + 24 * kSPitch + c0 + 2 * tg);
      };
#pragma unroll 1
      for (int pr = npairs - 1; pr >= 0; --pr) {
        const bool hasB = 2 * pr + 1 < nchunks;
#pragma unroll 1
        for (int c = 0; c < (hasB ? 2 : 1); ++c) {
          const int ch = 2 * pr + c;
          const int stage = (pos + c) % kSStages;
          const int ncols = min(kSChunk, D - ch * kSChunk);
          const float* Es = ring + size_t(stage) * STAGE;
          mbar_wait(&full[stage], ((pos + c) / kSStages) & 1);
#pragma unroll 1
          for (int c32 = warp * 32; c32 < ncols; c32 += NW * 32) {
            const bool whole = c32 + 32 <= ncols;
            float braw[KT][2];
            float2 e24 = make_float2(0.0f, 0.0f);
            load_b(Es, c32, braw, e24);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c0 = c32 + 8 * u;
              // round-to-nearest tf32 split on the integer pipe: hi = (b + 0x1000) & ~0x1fff, lo = b - hi
              uint32_t bhi[KT][2], blo[KT][2];
#pragma unroll
              for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  bhi[kt][h] = (__float_as_uint(braw[kt][h]) + 0x1000u) & 0xffffe000u;
                  blo[kt][h] = __float_as_uint(braw[kt][h] - __uint_as_float(bhi[kt][h]));
                }
              float accm[2][4], accc[2][4];
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                accm[mt][0] = m24[2 * mt] * e24.x; accm[mt][1] = m24[2 * mt] * e24.y;
                accm[mt][2] = m24[2 * mt + 1] * e24.x; accm[mt][3] = m24[2 * mt + 1] * e24.y;
                accc[mt][0] = accc[mt][1] = accc[mt][2] = accc[mt][3] = 0.0f;
              }
              if (u < 3) load_b(Es, c0 + 8, braw, e24);    // next 8 columns in flight behind this tile's MMAs
#pragma unroll
              for (int kt = 0; kt < KT; ++kt) {
                mma_tf32(accc[0], alo[0][kt], bhi[kt][0], bhi[kt][1]);
                mma_tf32(accc[1], alo[1][kt], bhi[kt][0], bhi[kt][1]);
                mma_tf32(accm[0], ahi[0][kt], bhi[kt][0], bhi[kt][1]);
                mma_tf32(accm[1], ahi[1][kt], bhi[kt][0], bhi[kt][1]);
                mma_tf32(accc[0], ahi[0][kt], blo[kt][0], blo[kt][1]);
                mma_tf32(accc[1], ahi[1][kt], blo[kt][0], blo[kt][1]);
              }
              {
                const bool cv = whole || c0 + 2 * tg < ncols;
                float* d = dbase + size_t(ch) * kSChunk + c0;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  stg_hint2_if(cv && stv[e], d + e * D8, accm[e >> 1][2 * (e & 1)] + accc[e >> 1][2 * (e & 1)],
                               accm[e >> 1][2 * (e & 1) + 1] + accc[e >> 1][2 * (e & 1) + 1], pol_drop);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
        }
        pos += hasB ? 2 : 1;
      }
    } } else if (need_bwd) {
      float* dE_t = demb + size_t(t) * S * D;
#pragma unroll 1
      for (int pr = npairs - 1; pr >= 0; --pr) {
        const int chA = 2 * pr, chB = 2 * pr + 1;
        const bool hasB = chB < nchunks;
        const int stA = pos % kSStages, stB = (pos + 1) % kSStages;
        const int nqA = min(kSChunk, D - chA * kSChunk) >> 2;
        const int nqB = hasB ? (min(kSChunk, D - chB * kSChunk) >> 2) : 0;
          mbar_wait(&full[stA], (pos / kSStages) & 1);
        if (hasB) mbar_wait(&full[stB], ((pos + 1) / kSStages) & 1);
        const int nq = nqA + nqB;
        for (int item = tid; item < 2 * nq; item += kSConsumers) {
          const int h = item / nq, c4 = item - h * nq;
          const float* ecol = (c4 < nqA) ? ring + size_t(stA) * STAGE + 4 * c4 : ring + size_t(stB) * STAGE + 4 * (c4 - nqA);
          float* dcol = dE_t + size_t(chA) * kSChunk + 4 * c4;
          if (h == 0) {
            float2 o[HR][2];
            bwd_tile<HR, kSPitch, 2 * HRP>(ecol, Mt, S, o);
#pragma unroll
            for (int r = 0; r < HR; ++r)
              if (r < S)
                stg_hint(reinterpret_cast<float4*>(dcol + size_t(r) * D), make_float4(o[r][0].x, o[r][0].y, o[r][1].x, o[r][1].y),
                         pol_drop);
          } else {
            float2 o[SG - HR][2];
            bwd_tile<SG - HR, kSPitch, 2 * HRP>(ecol, Mt + HRP, S, o);
#pragma unroll
            for (int r = 0; r < SG - HR; ++r)
              if (HR + r < S)
                stg_hint(reinterpret_cast<float4*>(dcol + size_t(HR + r) * D),
                         make_float4(o[r][0].x, o[r][0].y, o[r][1].x, o[r][1].y), pol_drop);
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&empty[stA]);
          if (hasB) mbar_arrive(&empty[stB]);
        }
        pos += hasB ? 2 : 1;
      }
    }
  }
  if (warp == 0) tup_publish_losses(done_counter, int(iter), T, loss_out, lane);
}

// ---------------------------------------------------------------------------------------------
template <int TS, int NW, int CH, bool kPackedGram = true, bool kMmaBwd = false>
static int stream_launch(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params& p,
                         float* per_tuple, float* demb, uint32_t* kept, float* loss, unsigned int* counter,
                         cudaStream_t stream) {
  auto kern = wms_stream_kernel<TS, NW, CH, kPackedGram, kMmaBwd>;
  constexpr size_t smem = SSmem<TS, CH, kMmaBwd>::bytes;
  static SmemAttrCache configured;                      // per device
  int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, &configured);
  if (rc_attr) return rc_attr;
  const int per_sm = (TS == 5 && NW <= 8) ? 2 : 1;
  int grid = num_sms() * per_sm;
  if (grid > T) grid = T;
  kern<<<grid, NW * 32 + 32, smem, stream>>>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

template <int TS>
static int stream_dispatch(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params& p,
                           float* per_tuple, float* demb, uint32_t* kept, float* loss, unsigned int* counter,
                           cudaStream_t stream) {
  // FFMA2 backward: 1: <16,512>, 2: <8,256> x 2 CTAs/SM, 3: <15,480>, 4: <7,224> x 2, 5: as 2 with a scalar-FFMA Gram.
  // Tensor-core backward: 6: <7,224> x 2 CTAs/SM (default for S <= 25).  Measured on B200, T = 4096, S = 25, D = 4096:
  // cfg 6 0.96 ms, cfg 2 1.04-1.08 ms (HBM reads 2.5 GB), cfg 4 1.12 ms, cfg 3 1.13 ms (1.98 GB), cfg 1 1.30 ms (1.89 GB):
  // one tuple pipeline per SM keeps the re-read in L2 but leaves the pipes idle during the scalar phase and the barriers.
  // Also measured and dropped: tensor-core backward with 6 warps x 192 columns x 5 stages 1.11 ms, with 8 warps x 256
  // columns at 96 registers (spills) 1.08 ms; a one-CTA-per-SM warp-specialised pipeline (Gram warps / weight warps,
  // 168 registers) 1.63 ms -- with this much register tile per warp, warps per SM are what hides the latencies.
  // Latency tweaks inside cfg 6 that did not pay: six instead of four accumulator chains in the backward 0.97 ms,
  // Gram operands software-pipelined one column quad ahead 1.02 ms; the backward on mma.sync.m16n8k16 f16 (hi/lo split
  // into fp16 with exact power-of-two scaling per tuple: 12 instead of 18 MMAs per 8 columns at the same instruction rate,
  // tools/ubench/mma_f16.cu) passed every parity test and ran 0.956 ms -- the backward is not bound by the tensor pipe:
  // the unit at its limit is the L1/shared-memory data pipe (71 % of peak over the whole kernel: Gram operand loads,
  // one wavefront per 32-byte sector of these accumulator-layout stores, fragment loads; profiles/r1_wms_l1_analysis.txt).  Trading registers for warps does not help either: the fp16
  // backward with 10 consumer warps x 160 columns x 5 stages at 80 registers (2 CTAs per SM) ran 1.24-1.26 ms.
  // cp.async.bulk.prefetch.L2 of the chunks 4 / 8 positions beyond the ring: 1.07 / 1.09 ms (the prefetched lines push
  // the tuples waiting for their re-read out of L2).
  // Round 2, the store path: the gradient block written IN PLACE over the descriptor block it came from (st.shared.v2,
  // 2 wavefronts per 256 bytes instead of 8 sectors) and the finished chunk stored by the producer warp with
  // cp.async.bulk shared -> global (25 rows x 896 bytes, one bulk group in flight, loads held back one position so the
  // warp never blocks on wait_group.read): every parity test passed, 1.16 ms.  By elimination (same build with pieces
  // removed): without the proxy fence 1.16, without the in-place writes 1.14, WITHOUT THE BULK STORES 0.925, with no
  // output at all 0.90 ms -- the bulk store engine is slow on 896-byte rows at a 944-byte pitch, and the whole direct
  // st.global path of this kernel costs only 0.055 ms: the stores are not what holds the kernel at 0.54 of HBM.  The
  // 0.90 ms is Gram + weights + tensor-core backward with the SM's two CTAs interleaved (HBM reads alone: 0.26 ms).
  // Forward only (need_bwd = false) 0.467 ms; with the Gram's FFMA2 removed (operand loads kept) 0.417; with the whole
  // Gram loop skipped 0.312 (HBM floor of the read 0.26) and forward + backward 0.830: the Gram costs 0.125-0.155 ms that
  // do not overlap with the streaming, the backward adds ~0.5 ms against 0.38 ms of its own HBM traffic (gradient
  // write + the half of the re-read that misses L2).
  const int cfg = knob_or(KNOB_WMS_STREAM_CFG, TS == 5 ? 6 : 2);
  if (cfg == 6) return stream_launch<TS, 7, 224, true, true>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
  // the wider register tiles (S > 25) need more registers than 16+ warps leave per thread
  if constexpr (TS == 5) {
    if (cfg == 5) return stream_launch<TS, 8, 256, false>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
    if (cfg == 4) return stream_launch<TS, 7, 224>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
    if (cfg == 3) return stream_launch<TS, 15, 480>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
    if (cfg == 1) return stream_launch<TS, 16, 512>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
  }
  return stream_launch<TS, 8, 256>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
}

// SCL_ERR_UNSUPPORTED: small batches go to the cluster kernels (more SMs per tuple).
int wms_stream_launch(const float* emb, const float* dist, int T, int S, int D, const scl_ms_params& p, float* loss,
                      float* per_tuple, float* demb, uint32_t* kept, unsigned int* counter, cudaStream_t stream) {
  const int mode = knob_or(KNOB_WMS_STREAM, -1);        // 0: never, 1: always, unset: large batches
  if (mode == 0) return SCL_ERR_UNSUPPORTED;
  if (S < 2 || S > 32 || D < 4 || (D & 3)) return SCL_ERR_UNSUPPORTED;
  if (mode != 1 && T < num_sms()) return SCL_ERR_UNSUPPORTED;
  if (S <= 25) return stream_dispatch<5>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
  if (S <= 30) return stream_dispatch<6>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
  return stream_dispatch<7>(emb, dist, T, S, D, p, per_tuple, demb, kept, loss, counter, stream);
}

}  // namespace scl
