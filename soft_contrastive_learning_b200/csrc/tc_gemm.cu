// tc_gemm.cu -- fp32-in / fp32-out GEMM on tcgen05 (kind::tf32), the contraction engine of the PCA projection
// (train/train.py:646-652), the flat-mode Gram matrix E E^T and its backward M E (model/losses.py:25, :94), and the
// NetVLAD contractions (model/nets.py:66-67).
//
//   C[M,N] = (A . B^T) * colscale[n]        A: M x K, B: N x K, each either K-major ([rows,K] row-major) or
//                                            MN-major ([K,rows] row-major: the transposed operand, no copy)
//
// Precision modes
//   x1  one kind::tf32 MMA per product: operands are read as fp32 and truncated to tf32 by the tensor core
//       (10-bit mantissa, relative error ~1e-3 per product; stated tolerance of callers: 2e-3 of the result norm);
//   x3  fp32-grade "3xTF32": hi = rn_tf32(x), lo = x - hi (exact in fp32; the tensor core truncates it to tf32),
//       D += hi_a hi_b + hi_a lo_b + lo_a hi_b.  Four "splitter" warps rewrite the TMA-landed fp32 tile in place as hi
//       and write the lo tile next to it (same swizzled offsets, so no layout math); neither ever touches HBM.  The
//       dropped lo_a lo_b term and the rounding of lo are each <= 2^-22 relative and unbiased.
//
// Persistent CTAs (one per SM) walk the 128 x BN output tiles (cta_group::1, UMMA 128 x BN x 8), K swept in 32-float
// (128-byte, SWIZZLE_128B) stages through an mbarrier ring that keeps running across tiles: warp 0 = TMA producer,
// warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue (tcgen05.ld 32x32b: one accumulator row per thread),
// warps 6-9 = splitters (x3 only).  The two TMEM accumulators alternate across flush groups AND tiles, so the loads and
// MMAs of tile i+1 run under the global stores of tile i -- what the short-K, many-tile NetVLAD contractions
// (K = 64, 10 240 tiles) are bound by.
#include <atomic>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace scl {

using namespace tc;

constexpr int kGBM = 128, kGBK = 32;                 // tile rows, floats per stage along K (= one 128-byte swizzle row)
constexpr uint32_t kGABytes = kGBM * kGBK * 4;       // 16 KB

template <int BN, bool kX3>
struct GCfg {
  static constexpr uint32_t kBBytes = BN * kGBK * 4;
  static constexpr uint32_t kHiBytes = kGABytes + kBBytes;
  static constexpr uint32_t kStageBytes = kX3 ? 2 * kHiBytes : kHiBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kThreads = kX3 ? 320 : 192;
  static constexpr uint32_t kTmemCols = 2 * BN;         // two accumulators: the MMAs fill one while the other is drained
  // The tensor core adds into its fp32 accumulator with truncation: the error grows linearly with the number of
  // accumulation steps (measured 2e-8 of the result per step -> 2e-4 at K = 32768).  The accumulator is therefore
  // flushed into registers (round-to-nearest adds) every kFlush stages: 2 stages = 24 MMA steps in 3xTF32 mode.
  static constexpr int kFlush = kX3 ? 2 : 32;
};

struct GSmemTail {
  uint64_t full[8], split[8], empty[8], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

template <int BN, bool kX3, bool kAMn, bool kBMn>
__global__ void __launch_bounds__(GCfg<BN, kX3>::kThreads, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, TcGemmArgs g) {
  using Cfg = GCfg<BN, kX3>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GSmemTail* tail = reinterpret_cast<GSmemTail*>(smem + size_t(kStages) * Cfg::kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_k = (g.K + kGBK - 1) / kGBK;
  const int ntiles = g.tiles_n * g.tiles_m * g.batch * g.split_k;
  // tile -> (n tile, m tile, batch index, K slice); consecutive tiles share their A rows
  struct Tile { int m0, n0, bz, k_begin, num_k; };
  auto decode = [&](int tile) {
    Tile t;
    const int bx = tile % g.tiles_n, r = tile / g.tiles_n, by = r % g.tiles_m, z = r / g.tiles_m;
    t.m0 = by * kGBM;
    t.n0 = bx * BN;
    t.bz = z / g.split_k;
    const int ks = z - t.bz * g.split_k;
    t.k_begin = (total_k * ks) / g.split_k;
    t.num_k = (total_k * (ks + 1)) / g.split_k - t.k_begin;
    return t;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (g.k1_stages < total_k) {
      prefetch_tmap(&tmA2);
      prefetch_tmap(&tmB2);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->split[s], 4);
      mbar_init(&tail->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->acc_full[b], 1);
      mbar_init(&tail->acc_empty[b], 4);              // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tail->tmem_base, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // every lane runs the loop, one elected lane issues: the operands of the single-thread instructions are then provably
    // warp-uniform and go straight to uniform registers (tc_common.cuh: elect_one)
    {
      uint32_t kcg = 0;                                 // ring position, running across tiles
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const Tile t = decode(tile);
      const int m0 = t.m0, n0 = t.n0, bz = t.bz, k_begin = t.k_begin;
      for (int kc = 0; kc < t.num_k; ++kc, ++kcg) {
        const int stage = kcg % kStages;
        mbar_wait(&tail->empty[stage], ((kcg / kStages) & 1) ^ 1);
        uint8_t* sa = smem + size_t(stage) * Cfg::kStageBytes;
        uint8_t* sb = sa + kGABytes;
        // K stages beyond k1_stages come from the second operand pair (K-concatenated problem)
        const int ka = k_begin + kc;
        const bool second = ka >= g.k1_stages;
        const CUtensorMap* mA = second ? &tmA2 : &tmA;
        const CUtensorMap* mB = second ? &tmB2 : &tmB;
        const int kk = (second ? ka - g.k1_stages : ka) * kGBK;
        const int bzb = (second ? g.b2_shared : g.b_shared) ? 0 : bz;
        if (elect_one()) {
          mbar_arrive_expect_tx(&tail->full[stage], Cfg::kHiBytes);
          if (kAMn) {
#pragma unroll
            for (int grp = 0; grp < kGBM / 32; ++grp) tma_load_3d(sa + grp * 4096, mA, &tail->full[stage], m0 + 32 * grp, kk, bz);
          } else {
            tma_load_3d(sa, mA, &tail->full[stage], kk, m0, bz);
          }
          if (kBMn) {
#pragma unroll
            for (int grp = 0; grp < BN / 32; ++grp) tma_load_3d(sb + grp * 4096, mB, &tail->full[stage], n0 + 32 * grp, kk, bzb);
          } else {
            tma_load_3d(sb, mB, &tail->full[stage], kk, n0, bzb);
          }
        }
        __syncwarp();
      }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      constexpr uint32_t idesc = make_idesc(kFmtTF32, kGBM, BN) | (kAMn ? (1u << 15) : 0u) | (kBMn ? (1u << 16) : 0u);
      // K advance of 8 tf32 inside a stage: K-major +32 bytes within the swizzle row, MN-major +1024 bytes (next atom)
      constexpr uint64_t stepA = kAMn ? (1024 >> 4) : (32 >> 4), stepB = kBMn ? (1024 >> 4) : (32 >> 4);
      uint32_t kcg = 0, grpg = 0;                       // ring position and flush-group count, running across tiles
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int num_k = decode(tile).num_k;
      for (int kc = 0; kc < num_k; ++kc, ++kcg) {
        const int stage = kcg % kStages;
        const int in_grp = kc % Cfg::kFlush;
        const uint32_t buf = grpg & 1;
        if (in_grp == 0) {
          mbar_wait(&tail->acc_empty[buf], ((grpg >> 1) & 1) ^ 1);      // the epilogue drained this accumulator
          tc_fence_after();
        }
        mbar_wait(kX3 ? &tail->split[stage] : &tail->full[stage], (kcg / kStages) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(buf) * BN;
        const uint32_t sa = smem_u32(smem + size_t(stage) * Cfg::kStageBytes);
        const uint32_t sb = sa + kGABytes;
        const uint64_t da = kAMn ? smem_desc_sw128_mn(sa) : smem_desc_sw128(sa);
        const uint64_t db = kBMn ? smem_desc_sw128_mn(sb) : smem_desc_sw128(sb);
        const bool last = in_grp == Cfg::kFlush - 1 || kc == num_k - 1;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kGBK / 8; ++k)
            mma_tf32_ss(d_tmem, da + stepA * k, db + stepB * k, idesc, (in_grp | k) != 0 ? 1u : 0u);
          if (kX3) {
            const uint64_t dal = kAMn ? smem_desc_sw128_mn(sa + Cfg::kHiBytes) : smem_desc_sw128(sa + Cfg::kHiBytes);
            const uint64_t dbl = kBMn ? smem_desc_sw128_mn(sb + Cfg::kHiBytes) : smem_desc_sw128(sb + Cfg::kHiBytes);
#pragma unroll
            for (int k = 0; k < kGBK / 8; ++k) {
              mma_tf32_ss(d_tmem, da + stepA * k, dbl + stepB * k, idesc, 1u);     // hi_a * lo_b
              mma_tf32_ss(d_tmem, dal + stepA * k, db + stepB * k, idesc, 1u);     // lo_a * hi_b
            }
          }
          mma_commit(&tail->empty[stage]);              // frees the stage when these MMAs retire
          if (last) mma_commit(&tail->acc_full[buf]);
        }
        __syncwarp();
        if (last) ++grpg;
      }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue: TMEM -> register accumulators (per flush group) -> global =====================
    const int lq = warp & 3;                            // TMEM lane quarter this warp may read
    uint32_t grpg = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const Tile t = decode(tile);
    const int n0 = t.n0, bz = t.bz;
    const int m = t.m0 + lq * 32 + lane;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
    const int ngroups = (t.num_k + Cfg::kFlush - 1) / Cfg::kFlush;
#pragma unroll 1
    for (int grp = 0; grp < ngroups; ++grp, ++grpg) {
      const uint32_t buf = grpg & 1;
      mbar_wait(&tail->acc_full[buf], (grpg >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(lq * 32) << 16) + uint32_t(buf) * BN;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->acc_empty[buf]);
    }
    if (m < g.M) {
      float* crow = g.C + size_t(bz) * g.sC + size_t(m) * g.ldc;
      const float rs = g.rowscale ? __ldg(g.rowscale + size_t(bz) * g.rowscale_stride + m) : 1.0f;
      // Whole tiles of a plain store go out as 32-byte (one full sector) stores: the accumulator layout is one row per
      // thread, so a warp's store touches 32 rows -- 16-byte pieces would write every sector in two halves.
      const bool wide = n0 + BN <= g.N && g.split_k == 1 && !g.accumulate && !g.colscale && (g.ldc & 7) == 0 &&
                        (g.sC & 7) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 31) == 0;
      if (wide) {
#pragma unroll
        for (int j8 = 0; j8 < BN / 8; ++j8) {
          const float* a8 = acc + 8 * j8;
          asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(crow + n0 + 8 * j8), "f"(a8[0] * rs),
                       "f"(a8[1] * rs), "f"(a8[2] * rs), "f"(a8[3] * rs), "f"(a8[4] * rs), "f"(a8[5] * rs), "f"(a8[6] * rs),
                       "f"(a8[7] * rs)
                       : "memory");
        }
      } else
#pragma unroll
      for (int j4 = 0; j4 < BN / 4; ++j4) {
        const int n = n0 + 4 * j4;
        if (n + 3 < g.N) {
          float4 o = make_float4(acc[4 * j4] * rs, acc[4 * j4 + 1] * rs, acc[4 * j4 + 2] * rs, acc[4 * j4 + 3] * rs);
          if (g.colscale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(g.colscale + n));
            o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w;
          }
          if (g.split_k > 1) {
            atomicAdd(crow + n, o.x); atomicAdd(crow + n + 1, o.y); atomicAdd(crow + n + 2, o.z); atomicAdd(crow + n + 3, o.w);
          } else {
            if (g.accumulate) {
              const float4 c0 = *reinterpret_cast<const float4*>(crow + n);
              o.x += c0.x; o.y += c0.y; o.z += c0.z; o.w += c0.w;
            }
            *reinterpret_cast<float4*>(crow + n) = o;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (n + u < g.N) {
              const float v = acc[4 * j4 + u] * rs * (g.colscale ? g.colscale[n + u] : 1.0f);
              if (g.split_k > 1) atomicAdd(crow + n + u, v);
              else crow[n + u] = g.accumulate ? crow[n + u] + v : v;
            }
        }
      }
    }
    }
  } else if (kX3) {
    // ===================== splitters: lo = x - tf32_trunc(x), same swizzled offsets =====================
    const int st = threadIdx.x - 192;                   // 0..127
    constexpr int kVec = int(Cfg::kHiBytes / 16);       // float4s in the hi part of a stage
    uint32_t nk_all = 0;                                // stages this CTA sweeps over all its tiles
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) nk_all += uint32_t(decode(tile).num_k);
    for (uint32_t kc = 0; kc < nk_all; ++kc) {
      const int stage = kc % kStages;
      mbar_wait(&tail->full[stage], (kc / kStages) & 1);
      float4* hi = reinterpret_cast<float4*>(smem + size_t(stage) * Cfg::kStageBytes);
      float4* lo = reinterpret_cast<float4*>(smem + size_t(stage) * Cfg::kStageBytes + Cfg::kHiBytes);
#pragma unroll 4
      for (int i = st; i < kVec; i += 128) {
        const float4 x = hi[i];
        float4 h, l;
        // lo = x - hi is exact in fp32 and is handed over unrounded: the tensor core drops its bits below tf32
        // (2^-22 of x, sign-symmetric like the residual itself)
        h.x = tf32_rn(x.x); l.x = x.x - h.x;
        h.y = tf32_rn(x.y); l.y = x.y - h.y;
        h.z = tf32_rn(x.z); l.z = x.z - h.z;
        h.w = tf32_rn(x.w); l.w = x.w - h.w;
        hi[i] = h;                                      // tf32-representable: whatever rounding the MMA applies is a no-op
        lo[i] = l;
      }
      fence_proxy_async();                              // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->split[stage]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
template <int BN, bool kX3, bool kAMn, bool kBMn>
static int tc_gemm_launch(const TcGemmDesc& d, cudaStream_t stream) {
  using Cfg = GCfg<BN, kX3>;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: dims {K, rows, batch}, box {32, tile rows, 1}; MN-major: dims {rows, K, batch}, box {32, 32, 1}
  const uint64_t nb = d.batch > 1 ? uint64_t(d.batch) : 1;
  const uint64_t pA = (nb > 1 ? uint64_t(d.sA) : uint64_t(d.lda) * (d.a_mn ? d.K : d.M)) * 4;
  const uint64_t pB = ((nb > 1 && d.sB) ? uint64_t(d.sB) : uint64_t(d.ldb) * (d.b_mn ? d.K : d.N)) * 4;
  const uint64_t nbB = (nb > 1 && d.sB == 0) ? 1 : nb;      // shared B: every batch reads slice 0
  if (kAMn) rc = make_tmap_3d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.A, uint64_t(d.M), uint64_t(d.K), nb, uint64_t(d.lda) * 4, pA, 32, 32, 1);
  else rc = make_tmap_3d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.A, uint64_t(d.K), uint64_t(d.M), nb, uint64_t(d.lda) * 4, pA, 32, kGBM, 0);
  if (rc) return rc;
  if (kBMn) rc = make_tmap_3d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.B, uint64_t(d.N), uint64_t(d.K), nbB, uint64_t(d.ldb) * 4, pB, 32, 32, 1);
  else rc = make_tmap_3d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.B, uint64_t(d.K), uint64_t(d.N), nbB, uint64_t(d.ldb) * 4, pB, 32, BN, 0);
  if (rc) return rc;
  // optional second operand pair, K-concatenated behind the first
  CUtensorMap tmA2 = tmA, tmB2 = tmB;
  const bool dual = d.A2 != nullptr;
  if (dual) {
    if (!d.B2 || d.K2 < 1 || (d.K % kGBK) != 0 || d.split_k == 2 || !aligned16(d.A2) || !aligned16(d.B2)) return SCL_ERR_BAD_ARG;
    const uint64_t pA2 = (nb > 1 ? uint64_t(d.sA) : uint64_t(d.lda) * (d.a_mn ? d.K2 : d.M)) * 4;
    const uint64_t pB2 = ((nb > 1 && d.sB2) ? uint64_t(d.sB2) : uint64_t(d.ldb) * (d.b_mn ? d.K2 : d.N)) * 4;
    const uint64_t nbB2 = (nb > 1 && d.sB2 == 0) ? 1 : nb;
    if (kAMn) rc = make_tmap_3d(&tmA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.A2, uint64_t(d.M), uint64_t(d.K2), nb, uint64_t(d.lda) * 4, pA2, 32, 32, 1);
    else rc = make_tmap_3d(&tmA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.A2, uint64_t(d.K2), uint64_t(d.M), nb, uint64_t(d.lda) * 4, pA2, 32, kGBM, 0);
    if (rc) return rc;
    if (kBMn) rc = make_tmap_3d(&tmB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.B2, uint64_t(d.N), uint64_t(d.K2), nbB2, uint64_t(d.ldb) * 4, pB2, 32, 32, 1);
    else rc = make_tmap_3d(&tmB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.B2, uint64_t(d.K2), uint64_t(d.N), nbB2, uint64_t(d.ldb) * 4, pB2, 32, BN, 0);
    if (rc) return rc;
  }
  auto kern = tc_gemm_kernel<BN, kX3, kAMn, kBMn>;
  const size_t smem = 1024 + size_t(Cfg::kStages) * Cfg::kStageBytes + sizeof(GSmemTail);
  static SmemAttrCache configured;                     // per device
  rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, &configured);
  if (rc) return rc;
  TcGemmArgs g;
  g.M = d.M; g.N = d.N; g.K = d.K + (dual ? d.K2 : 0); g.colscale = d.colscale; g.C = d.C; g.ldc = d.ldc;
  g.k1_stages = dual ? d.K / kGBK : 0x7fffffff;
  g.b2_shared = (dual && nb > 1 && d.sB2 == 0) ? 1 : 0;
  g.sC = d.sC; g.rowscale = d.rowscale; g.rowscale_stride = d.rowscale_stride; g.accumulate = d.accumulate;
  g.batch = int(nb);
  g.b_shared = (nb > 1 && d.sB == 0) ? 1 : 0;
  g.split_k = (d.split_k == 2 && (d.K + kGBK - 1) / kGBK >= 2 && !d.accumulate) ? 2 : 1;
  g.tiles_n = (d.N + BN - 1) / BN;
  g.tiles_m = (d.M + kGBM - 1) / kGBM;
  const long long ntiles = (long long)g.tiles_n * g.tiles_m * g.batch * g.split_k;
  if (ntiles > 0x7fffffffLL) return SCL_ERR_BAD_SHAPE;
  const unsigned grid = unsigned(ntiles < num_sms() ? ntiles : num_sms());      // persistent: one CTA per SM
  kern<<<grid, Cfg::kThreads, smem, stream>>>(tmA, tmB, tmA2, tmB2, g);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

template <int BN, bool kX3>
static int tc_gemm_major(const TcGemmDesc& d, cudaStream_t stream) {
  if (d.a_mn) {
    if (d.b_mn) return tc_gemm_launch<BN, kX3, true, true>(d, stream);
    return tc_gemm_launch<BN, kX3, true, false>(d, stream);
  }
  if (d.b_mn) return tc_gemm_launch<BN, kX3, false, true>(d, stream);
  return tc_gemm_launch<BN, kX3, false, false>(d, stream);
}

static std::atomic<int> g_gemm_precision{0};
int tc_gemm_precision() { return g_gemm_precision.load(std::memory_order_relaxed); }

// Shapes: lda/ldb/ldc multiples of 4 floats, pointers 16-byte aligned, N multiple of 4.
int tc_gemm(const TcGemmDesc& d, cudaStream_t stream) {
  if (!d.A || !d.B || !d.C || d.M < 1 || d.N < 1 || d.K < 1) return SCL_ERR_BAD_ARG;
  if ((d.lda & 3) || (d.ldb & 3) || (d.ldc & 3) || !aligned16(d.A) || !aligned16(d.B) || !aligned16(d.C)) return SCL_ERR_ALIGN;
  const bool x3 = d.precision == 0;
  // narrow tiles for narrow outputs; otherwise 128 x 128 (x3 keeps hi+lo of both operands per stage)
  if (d.N <= 64) return x3 ? tc_gemm_major<64, true>(d, stream) : tc_gemm_major<64, false>(d, stream);
  return x3 ? tc_gemm_major<128, true>(d, stream) : tc_gemm_major<128, false>(d, stream);
}

}  // namespace scl

// precision of every tensor-core contraction behind the C ABI: 0 = fp32-grade (3xTF32, default), 1 = single TF32 pass
extern "C" int scl_set_gemm_precision(int mode) {
  if (mode != 0 && mode != 1) return SCL_ERR_BAD_ARG;
  scl::g_gemm_precision.store(mode, std::memory_order_relaxed);
  return SCL_OK;
}
extern "C" int scl_get_gemm_precision(void) { return scl::tc_gemm_precision(); }

// The contraction engine itself (tests, and callers that want it directly):
//   C[M,N] = A . B^T (* colscale[n]),  a_mn / b_mn select the transposed (MN-major) storage of an operand.
extern "C" int scl_gemm_tf32(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                             int a_mn, int b_mn, const float* colscale, int precision, scl_stream_t stream) {
  int rc = scl::check_device();
  if (rc) return rc;
  scl::TcGemmDesc d = {};
  d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ldc = ldc;
  d.a_mn = a_mn != 0; d.b_mn = b_mn != 0; d.colscale = colscale; d.precision = precision;
  return scl::tc_gemm(d, static_cast<cudaStream_t>(stream));
}
