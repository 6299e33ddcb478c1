// netvlad_dx.cu -- second half of the NetVLAD backward: the gradient of the conv5 maps.
//
//   dxh[p,c] = sum_k a[p,k] dV[b][c,k] + sum_k ds[p,k] W[c,k]          (aggregation path + assignment path)
//   dx [p,c] = inv[p] * dxh[p,c] - rb[p] * x[p,c]                      (backward of tf.nn.l2_normalize, nets.py:66;
//                                                                       rb = inv^2 (xh . dxh), from netvlad_fused.cu)
//
// The generic path ran this as a 3xTF32 contraction over the concatenated K = 64 + 64 (tc_gemm.cu) that wrote dxh, and a
// second kernel that re-read x and dxh and rewrote dx: x once, dx written, read and written again, with a row-per-thread
// epilogue whose 32-byte requests bound the whole contraction.  Here every operand arrives already split into fp16
// hi / lo halves -- a and ds from the epilogues of the fused forward / backward kernels, dV[b] and W from their small
// split kernels -- so there are no splitter warps, the tensor core runs at the f16 rate, and the epilogue works on a
// shared-memory tile:  TMA lands the x tile, one thread per position combines it in place with its two accumulator rows
// (the two terms carry different power-of-two scales, so they keep separate accumulators), and TMA stores the tile.
// x is read once, dx written once, all of it in full 128-byte lines.
//
// Persistent CTAs; a unit is (image, 128-position tile): its A operand (a | ds, hi and lo: 64 KB) is loaded once and
// serves four 128-channel tiles.  B (dV[b] | W rows of the channel tile, hi and lo) streams through two 32 KB slots.
// The x / dx tile moves in four 32-channel quarters through a ring of 16 KB slots.  TMEM: 2 x (acc1 | acc2) = 512 columns,
// so the epilogue of a channel tile overlaps the MMAs of the next.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2-5 = epilogue.
#include <cuda_fp16.h>

#include "netvlad_fused.cuh"
#include "tc_common.cuh"

namespace scl {

using namespace tc;

constexpr uint32_t kDA = 16384;                        // one fp16 operand tile: 128 rows x 64 k
constexpr uint32_t kDOffA = 0;                         // a hi | a lo | ds hi | ds lo
constexpr uint32_t kDOffB = 4 * kDA;                   // two slots of (hi | lo) = 32 KB each
constexpr uint32_t kDOffX = kDOffB + 4 * kDA;          // x / dx quarters: kDXSlots x 16 KB
constexpr int kDXSlots = 4;
constexpr uint32_t kDOffTail = kDOffX + kDXSlots * kDA;
constexpr int kDThreads = 192;

struct DSmemTail {
  uint64_t a_full, a_empty, s_full, s_empty, b_full[2], b_empty[2], x_full[kDXSlots], x_empty[kDXSlots], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct NvDxArgs {
  int B, HW, C, tpi, units;
  const float* inv;      // [B*HW]
  const float* rb;       // [B*HW]
  const float* dvun;     // [B] 2^-e of dV[b]
  const float* dsscale;  // [B] scale of ds
  const float* wun;      // 2^-e of W
};

__global__ void __launch_bounds__(kDThreads, 1)
    nv_dx_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                 const __grid_constant__ CUtensorMap tmSh, const __grid_constant__ CUtensorMap tmSl,
                 const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                 const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDx, NvDxArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  DSmemTail* tail = reinterpret_cast<DSmemTail*>(smem + kDOffTail);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = int(gridDim.x);
  const int u0 = int((long long)blockIdx.x * g.units / G), u1 = int((long long)(blockIdx.x + 1) * g.units / G);
  const int NT = g.C / 128;                            // channel tiles per unit

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmSh); prefetch_tmap(&tmSl);
    prefetch_tmap(&tmVh); prefetch_tmap(&tmVl); prefetch_tmap(&tmWh); prefetch_tmap(&tmWl);
    prefetch_tmap(&tmX); prefetch_tmap(&tmDx);
    mbar_init(&tail->a_full, 1);
    mbar_init(&tail->a_empty, 1);
    mbar_init(&tail->s_full, 1);
    mbar_init(&tail->s_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tail->b_full[i], 1);
      mbar_init(&tail->b_empty[i], 1);
      mbar_init(&tail->acc_full[i], 1);
      mbar_init(&tail->acc_empty[i], 4);
    }
    for (int i = 0; i < kDXSlots; ++i) {
      mbar_init(&tail->x_full[i], 1);
      mbar_init(&tail->x_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tail->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t bq = 0, xq = 0;                           // B slot uses, x quarter uses
    for (int u = u0; u < u1; ++u) {
      const int b = u / g.tpi, pos0 = (u - b * g.tpi) * 128, ul = u - u0;
      // A operand of the unit: a (term 1) and ds (term 2) have their own barriers, each is free again as soon as its last
      // MMA of the previous unit has retired
      mbar_wait(&tail->a_empty, (ul & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&tail->a_full, 2 * kDA);
        tma_load_3d(smem + kDOffA, &tmAh, &tail->a_full, 0, pos0, b);
        tma_load_3d(smem + kDOffA + kDA, &tmAl, &tail->a_full, 0, pos0, b);
      }
      __syncwarp();
      mbar_wait(&tail->s_empty, (ul & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&tail->s_full, 2 * kDA);
        tma_load_3d(smem + kDOffA + 2 * kDA, &tmSh, &tail->s_full, 0, pos0, b);
        tma_load_3d(smem + kDOffA + 3 * kDA, &tmSl, &tail->s_full, 0, pos0, b);
      }
      __syncwarp();
      for (int n = 0; n < NT; ++n) {
        for (int term = 0; term < 2; ++term, ++bq) {   // B slot: dV[b] rows, then W rows of this channel tile
          const int slot = bq & 1;
          mbar_wait(&tail->b_empty[slot], ((bq >> 1) & 1) ^ 1);
          uint8_t* sb = smem + kDOffB + size_t(slot) * 2 * kDA;
          if (elect_one()) {
            mbar_arrive_expect_tx(&tail->b_full[slot], 2 * kDA);
            if (term == 0) {
              tma_load_3d(sb, &tmVh, &tail->b_full[slot], 0, n * 128, b);
              tma_load_3d(sb + kDA, &tmVl, &tail->b_full[slot], 0, n * 128, b);
            } else {
              tma_load_3d(sb, &tmWh, &tail->b_full[slot], 0, n * 128, 0);
              tma_load_3d(sb + kDA, &tmWl, &tail->b_full[slot], 0, n * 128, 0);
            }
          }
          __syncwarp();
        }
        for (int q = 0; q < 4; ++q, ++xq) {            // the x tile, 32 channels at a time
          const int slot = xq % kDXSlots;
          mbar_wait(&tail->x_empty[slot], ((xq / kDXSlots) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&tail->x_full[slot], kDA);
            tma_load_3d(smem + kDOffX + size_t(slot) * kDA, &tmX, &tail->x_full[slot], n * 128 + 32 * q, pos0, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(kFmtF16, 128, 128);
    uint32_t bq = 0, tq = 0;                           // B slot uses, channel tiles done
    for (int u = u0; u < u1; ++u) {
      const int ul = u - u0;
      for (int n = 0; n < NT; ++n, ++tq) {
        const uint32_t buf = tq & 1;
        mbar_wait(&tail->acc_empty[buf], ((tq >> 1) & 1) ^ 1);
        for (int term = 0; term < 2; ++term, ++bq) {
          const int slot = bq & 1;
          if (n == 0) mbar_wait(term == 0 ? &tail->a_full : &tail->s_full, ul & 1);
          mbar_wait(&tail->b_full[slot], (bq >> 1) & 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * 256 + term * 128;
          const uint32_t sa = smem_u32(smem + kDOffA + size_t(term) * 2 * kDA);
          const uint32_t sb = smem_u32(smem + kDOffB + size_t(slot) * 2 * kDA);
          const uint64_t dah = smem_desc_sw128(sa), dal = smem_desc_sw128(sa + kDA);
          const uint64_t dbh = smem_desc_sw128(sb), dbl = smem_desc_sw128(sb + kDA);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_f16_ss(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, k != 0 ? 1u : 0u);
              mma_f16_ss(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1u);
              mma_f16_ss(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, 1u);
            }
            mma_commit(&tail->b_empty[slot]);
            if (n == NT - 1) mma_commit(term == 0 ? &tail->a_empty : &tail->s_empty);
            if (term == 1) mma_commit(&tail->acc_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue: one thread per position =====================
    const int lq = warp & 3;
    const int row = lq * 32 + lane;
    const int sw = row & 7;
    uint32_t tq = 0, xq = 0;
    for (int u = u0; u < u1; ++u) {
      const int b = u / g.tpi, pos0 = (u - b * g.tpi) * 128;
      const int p = pos0 + row;
      const bool valid = p < g.HW;
      const float iv = valid ? __ldg(g.inv + size_t(b) * g.HW + p) : 0.0f;
      const float rbv = valid ? __ldg(g.rb + size_t(b) * g.HW + p) : 0.0f;
      const float u1s = iv * __ldg(g.dvun + b) * (1.0f / 16384.0f);      // a was scaled by 2^14, dV[b] by 1 / dvun[b]
      const float u2s = iv * __ldg(g.wun) / __ldg(g.dsscale + b);         // ds by dsscale[b], W by 1 / wun
      for (int n = 0; n < NT; ++n, ++tq) {
        const uint32_t buf = tq & 1;
        mbar_wait(&tail->acc_full[buf], (tq >> 1) & 1);
        tc_fence_after();
        const uint32_t t1 = tmem_base + (uint32_t(lq * 32) << 16) + buf * 256;
        for (int q = 0; q < 4; ++q, ++xq) {
          const int slot = xq % kDXSlots;
          uint32_t v1[32], v2[32];
          tmem_ld_32x32(t1 + 32 * q, v1);
          tmem_ld_32x32(t1 + 128 + 32 * q, v2);
          mbar_wait(&tail->x_full[slot], (xq / kDXSlots) & 1);
          tmem_ld_wait();
          uint8_t* xrow = smem + kDOffX + size_t(slot) * kDA + row * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {                 // chunk c of this row sits at c ^ (row & 7)
            float4* px = reinterpret_cast<float4*>(xrow + ((c ^ sw) << 4));
            const float4 xv = *px;
            float4 o;
            o.x = fmaf(-rbv, xv.x, fmaf(__uint_as_float(v1[4 * c]), u1s, __uint_as_float(v2[4 * c]) * u2s));
            o.y = fmaf(-rbv, xv.y, fmaf(__uint_as_float(v1[4 * c + 1]), u1s, __uint_as_float(v2[4 * c + 1]) * u2s));
            o.z = fmaf(-rbv, xv.z, fmaf(__uint_as_float(v1[4 * c + 2]), u1s, __uint_as_float(v2[4 * c + 2]) * u2s));
            o.w = fmaf(-rbv, xv.w, fmaf(__uint_as_float(v1[4 * c + 3]), u1s, __uint_as_float(v2[4 * c + 3]) * u2s));
            *px = o;
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");  // the whole quarter is written
          if (warp == 2 && elect_one()) {
            tma_store_3d(&tmDx, smem + kDOffX + size_t(slot) * kDA, n * 128 + 32 * q, pos0, b);
            tma_store_commit();
            // two stores stay in flight; the quarter stored two steps ago has been read out of its slot by now
            tma_store_wait_read<2>();
            if (xq >= 2) mbar_arrive(&tail->x_empty[(xq - 2) % kDXSlots]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->acc_empty[buf]);
      }
    }
    if (warp == 2 && elect_one()) {
      tma_store_wait_read<0>();
      for (uint32_t t = xq >= 2 ? xq - 2 : 0; t < xq; ++t) mbar_arrive(&tail->x_empty[t % kDXSlots]);
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dx [B,HW,C] from the fp16 hi / lo operands (see the header of this file).  a_hl / ds_hl: [B*HW,64] hi and lo,
// dv_hl: [B,C,64] hi and lo, w_hl: [C,64] hi and lo.
int nv_dx(const float* x, const __half* a_hi, const __half* a_lo, const __half* ds_hi, const __half* ds_lo, const __half* dv_hi,
          const __half* dv_lo, const __half* w_hi, const __half* w_lo, const float* inv, const float* rb, const float* dvun,
          const float* dsscale, const float* wun, int B, int HW, int C, float* dx, cudaStream_t stream) {
  CUtensorMap tmAh, tmAl, tmSh, tmSl, tmVh, tmVl, tmWh, tmWl, tmX, tmDx;
  int rc;
  const auto h = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  if ((rc = make_tmap_3d(&tmAh, h, a_hi, 64, uint64_t(HW), uint64_t(B), 128, uint64_t(HW) * 128, 64, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmAl, h, a_lo, 64, uint64_t(HW), uint64_t(B), 128, uint64_t(HW) * 128, 64, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmSh, h, ds_hi, 64, uint64_t(HW), uint64_t(B), 128, uint64_t(HW) * 128, 64, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmSl, h, ds_lo, 64, uint64_t(HW), uint64_t(B), 128, uint64_t(HW) * 128, 64, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmVh, h, dv_hi, 64, uint64_t(C), uint64_t(B), 128, uint64_t(C) * 128, 64, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmVl, h, dv_lo, 64, uint64_t(C), uint64_t(B), 128, uint64_t(C) * 128, 64, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmWh, h, w_hi, 64, uint64_t(C), 1, 128, uint64_t(C) * 128, 64, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmWl, h, w_lo, 64, uint64_t(C), 1, 128, uint64_t(C) * 128, 64, 128, 0))) return rc;
  const uint64_t pitchX = uint64_t(C) * 4, batchX = uint64_t(HW) * C * 4;
  if ((rc = make_tmap_3d(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, uint64_t(C), uint64_t(HW), uint64_t(B), pitchX, batchX, 32, 128, 0))) return rc;
  if ((rc = make_tmap_3d(&tmDx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dx, uint64_t(C), uint64_t(HW), uint64_t(B), pitchX, batchX, 32, 128, 0))) return rc;
  NvDxArgs g;
  g.B = B; g.HW = HW; g.C = C; g.tpi = (HW + 127) / 128;
  const long long units = (long long)B * g.tpi;
  g.units = int(units);
  g.inv = inv; g.rb = rb; g.dvun = dvun; g.dsscale = dsscale; g.wun = wun;
  const int G = int(units < num_sms() ? units : num_sms());
  const size_t smem = 1024 + size_t(kDOffTail) + sizeof(DSmemTail);
  static SmemAttrCache configured;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(nv_dx_kernel), smem, &configured))) return rc;
  nv_dx_kernel<<<G, kDThreads, smem, stream>>>(tmAh, tmAl, tmSh, tmSl, tmVh, tmVl, tmWh, tmWl, tmX, tmDx, g);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

}  // namespace scl
