// knn_tc.cu -- R1 candidate pass: fp16 distance GEMM on tcgen05 with a fused per-row top-k' filter.
//
// Replaces the O(Q*R*D) part of KDTree(ref).query(query, k) (/root/reference/evaluation/top-n.py:103-106).
//   scores(q, r) = |r|^2 - 2 q.r        (|q|^2 is constant per query and does not change the order)
// A = queries (fp16, [Q,Dp] K-major), B = database shard (fp16 shadow, [R,Dp] K-major), both brought in by TMA as
// 128-byte-swizzled 64-column boxes; 128x256 accumulator tiles live in TMEM (two of them, so the epilogue of tile i
// overlaps the MMAs of tile i+1); one elected thread issues tcgen05.mma.  The epilogue never writes the score
// matrix: thread `row` of the 128 epilogue threads owns query row `row` of the tile (tcgen05.ld 32x32b hands every
// thread one TMEM lane), compares its 256 scores against the row's running threshold and appends the survivors to the
// row's candidate list in global memory; when a list fills up the warp prunes it cooperatively to the k' best and
// tightens the threshold.  Lists are exact for the fp16 scores: everything that was ever rejected or pruned scored
// >= the final threshold, which is what knn.cu's exactness certificate relies on.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#include <cuda_fp16.h>

#include "knn_internal.cuh"
#include "tc_common.cuh"

namespace scl {

using namespace tc;

constexpr int kBM = 128, kBN = 256, kBK = 64;          // CTA tile; kBK fp16 = 128 bytes = one swizzle row
constexpr int kStages = 4;
constexpr int kUmmaK = 16;
constexpr int kTcThreads = 192;
constexpr uint32_t kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2;
constexpr uint32_t kStageBytes = kABytes + kBBytes;     // 48 KB
constexpr uint32_t kTmemCols = 512;                     // 2 accumulators x 256 fp32 columns

struct TcSmemTail {
  float rn[2][kBN];
  unsigned long long prune_keys[4][kCandCap];
  float prune_thr[4];
  uint64_t full[kStages], empty[kStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};
constexpr size_t kTcSmemBytes = 1024 + size_t(kStages) * kStageBytes + sizeof(TcSmemTail);

__device__ __forceinline__ unsigned long long cand_key(float s, uint32_t idx) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);       // order-preserving map float -> uint
  return (static_cast<unsigned long long>(u) << 32) | idx;
}

// Warp-cooperative prune of the candidate list of lane `L`'s row to its kKeep best entries (sorted ascending).
__device__ __forceinline__ void prune_row(int L, int lane, float* __restrict__ cs, uint32_t* __restrict__ ci,
                                          size_t base, int& cnt, float& thr, unsigned long long* keys, float* thr_slot) {
  const int n = __shfl_sync(0xffffffffu, cnt, L);
  const unsigned long long b64 = static_cast<unsigned long long>(base);
  const size_t bL = static_cast<size_t>(__shfl_sync(0xffffffffu, b64, L));
  __syncwarp();
  float ms[kCandCap / 32];
  uint32_t mi[kCandCap / 32];
#pragma unroll
  for (int u = 0; u < kCandCap / 32; ++u) {
    const int e = lane + 32 * u;
    if (e < n) {
      ms[u] = __ldcg(cs + bL + e);
      mi[u] = __ldcg(ci + bL + e);
      keys[e] = cand_key(ms[u], mi[u]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < kCandCap / 32; ++u) {
    const int e = lane + 32 * u;
    if (e < n) {
      const unsigned long long mine = keys[e];
      int rank = 0;
      for (int x = 0; x < n; ++x) rank += keys[x] < mine ? 1 : 0;
      if (rank < kKeep) {
        cs[bL + rank] = ms[u];
        ci[bL + rank] = mi[u];
        if (rank == kKeep - 1) *thr_slot = ms[u];
      }
    }
  }
  __syncwarp();
  if (lane == L) {
    cnt = kKeep;
    thr = *thr_slot;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kTcThreads, 1) knn_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(smem + size_t(kStages) * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], 4);     // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tail->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  const int num_k = a.Dp / kBK;
  const int num_items = a.num_m_blocks * a.NR;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int range = item / a.num_m_blocks, mb = item - range * a.num_m_blocks;
        const int t0 = range * a.tiles_per_range, t1 = min(a.num_n_tiles, t0 + a.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          for (int kc = 0; kc < num_k; ++kc) {
            mbar_wait(&tail->empty[stage], phase ^ 1);
            uint8_t* sa = smem + size_t(stage) * kStageBytes;
            mbar_arrive_expect_tx(&tail->full[stage], kStageBytes);
            tma_load_2d(sa, &tmA, &tail->full[stage], kc * kBK, mb * kBM);
            tma_load_2d(sa + kABytes, &tmB, &tail->full[stage], kc * kBK, t * kBN);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtF16, kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_count = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int range = item / a.num_m_blocks;
        const int t0 = range * a.tiles_per_range, t1 = min(a.num_n_tiles, t0 + a.tiles_per_range);
        for (int t = t0; t < t1; ++t, ++tile_count) {
          const uint32_t buf = tile_count & 1, use = tile_count >> 1;
          mbar_wait(&tail->tmem_empty[buf], (use & 1) ^ 1);       // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kBN;
          for (int kc = 0; kc < num_k; ++kc) {
            mbar_wait(&tail->full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + size_t(stage) * kStageBytes);
            const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 16 fp16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
              mma_f16_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
            }
            mma_commit(&tail->empty[stage]);                      // frees the smem slot when these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          mma_commit(&tail->tmem_full[buf]);                      // accumulator complete
        }
      }
    }
  } else {
    // ===================== epilogue: fused top-k' filter =====================
    const int ew = warp - 2;                 // 0..3, index into per-warp scratch
    const int lq = warp & 3;                 // TMEM lane quarter this warp may read
    const int et = ew * 32 + lane;           // 0..127 epilogue thread id
    unsigned long long* keys = tail->prune_keys[ew];
    float* thr_slot = &tail->prune_thr[ew];
    uint32_t tile_count = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int range = item / a.num_m_blocks, mb = item - range * a.num_m_blocks;
      const int t0 = range * a.tiles_per_range, t1 = min(a.num_n_tiles, t0 + a.tiles_per_range);
      const int q = mb * kBM + lq * 32 + lane;
      const bool valid = q < a.Q;
      const float qmul = valid ? a.qmul[q] : 0.0f;
      const size_t base = (size_t(valid ? q : 0) * a.NR + range) * kCandCap;
      float thr = INFINITY;
      int cnt = 0;
      for (int t = t0; t < t1; ++t, ++tile_count) {
        const uint32_t buf = tile_count & 1, use = tile_count >> 1;
        const int n0 = t * kBN;
        // stage the |r|^2 of this tile's 256 rows (rows beyond the shard score +inf)
        {
          const int r0 = n0 + et, r1 = n0 + 128 + et;
          tail->rn[buf][et] = r0 < a.R ? __ldg(a.rn + r0) : INFINITY;
          tail->rn[buf][128 + et] = r1 < a.R ? __ldg(a.rn + r1) : INFINITY;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tail->tmem_full[buf], use & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(lq * 32) << 16) + buf * kBN;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          const float* rn = &tail->rn[buf][c * 32];
          if (a.dbg_scores != nullptr && valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c * 32 + j;
              if (n < a.R) a.dbg_scores[size_t(q) * a.R + n] = fmaf(__uint_as_float(v[j]), qmul, rn[j]);
            }
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = fmaf(__uint_as_float(v[j]), qmul, rn[j]);
              if (s < thr) {
                a.cand_s[base + cnt] = s;
                a.cand_i[base + cnt] = uint32_t(n0 + c * 32 + j);
                ++cnt;
              }
            }
          }
          // keep at least 32 free slots before the next chunk
          unsigned need = __ballot_sync(0xffffffffu, cnt > kCandCap - 32);
          while (need) {
            const int L = __ffs(need) - 1;
            prune_row(L, lane, a.cand_s, a.cand_i, base, cnt, thr, keys, thr_slot);
            need &= need - 1;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->tmem_empty[buf]);
      }
      // end of the item: bring every list down to <= k' entries and publish its length
      unsigned need = __ballot_sync(0xffffffffu, cnt > kKeep);
      while (need) {
        const int L = __ffs(need) - 1;
        prune_row(L, lane, a.cand_s, a.cand_i, base, cnt, thr, keys, thr_slot);
        need &= need - 1;
      }
      if (valid) a.cand_cnt[size_t(q) * a.NR + range] = cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SCL_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) {
      set_last_error("cuTensorMapEncodeTiled entry point", cudaErrorUnknown);
      return SCL_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  (void)elem_bytes;
  CUresult r = fn(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    set_last_error(msg, cudaErrorInvalidValue);
    return SCL_ERR_CUDA;
  }
  return SCL_OK;
}

int knn_tc_launch(const TcArgs& a, const void* qh, const void* dbh, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, qh, uint64_t(a.Dp), uint64_t(a.Q),
                        uint64_t(a.Dp) * 2, kBK, kBM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dbh, uint64_t(a.Dp), uint64_t(a.R), uint64_t(a.Dp) * 2,
                    kBK, kBN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    SCL_CUDA_TRY(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTcSmemBytes)));
    configured = true;
  }
  const int items = a.num_m_blocks * a.NR;
  const int grid = items < num_sms() ? items : num_sms();
  knn_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(tmA, tmB, a);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

void knn_tc_tiling(int Q, int64_t R, int* num_m_blocks, int* num_n_tiles, int* NR, int* tiles_per_range) {
  const int mb = (Q + kBM - 1) / kBM;
  const int nt = int((R + kBN - 1) / kBN);
  const int sms = num_sms();
  // choose the number of database ranges so that (query blocks x ranges) fills whole waves of SMs
  int best = 1;
  double best_eff = -1.0;
  const int max_nr = nt < 64 ? nt : 64;
  for (int nr = 1; nr <= max_nr; ++nr) {
    const long long items = 1ll * mb * nr;
    const long long waves = (items + sms - 1) / sms;
    const int tpr = (nt + nr - 1) / nr;
    // time ~ waves * tiles_per_range; efficiency relative to perfect balance
    const double eff = double(1ll * mb * nt) / (double(waves) * sms * tpr);
    if (eff > best_eff + 0.005) { best_eff = eff; best = nr; }
  }
  const char* env = getenv("SCL_KNN_RANGES");
  if (env && atoi(env) >= 1 && atoi(env) <= max_nr) best = atoi(env);
  *num_m_blocks = mb;
  *num_n_tiles = nt;
  *NR = best;
  *tiles_per_range = (nt + best - 1) / best;
}

}  // namespace scl
