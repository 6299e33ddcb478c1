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
#include <atomic>

#include "tc_common.cuh"

namespace scl {

using namespace tc;

constexpr int kBM = 128, kBN = 256, kBK = 64;          // per-CTA accumulator tile; kBK fp16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kTcThreads = 192;
constexpr int kMaxStages = 6;
constexpr uint32_t kABytes = kBM * kBK * 2;             // 16 KB: this CTA's 128 query rows
constexpr uint32_t kTmemCols = 512;                     // 2 accumulators x 256 fp32 columns

// kPair = false: one CTA per 128x256 tile (cta_group::1), the CTA loads the whole 256-row B tile.
// kPair = true : a CTA pair computes a 256x256 tile (cta_group::2, UMMA M = 256); each CTA loads its own 128 query
//                rows and HALF of the B tile, the tensor cores read the other half from the peer's shared memory.
//                One third less L2->SM traffic per flop and two more pipeline stages.
// kNSub = 2   : the CTA keeps TWO 256-column accumulators (all 512 TMEM columns) alive for one K sweep, so each
//                query chunk fetched from L2 is used against 512 database rows instead of 256: another quarter less
//                L2->SM traffic per flop.  The accumulators are then single-buffered; with K = 4096 a tile computes for
//                ~60 us and drains in ~2 us, and TMA keeps prefetching the next tile's stages meanwhile.
template <bool kPair, int kNSub>
struct TcCfg {
  static constexpr uint32_t kBRows = kPair ? kBN / 2 : kBN;            // rows per B sub-tile held by this CTA
  static constexpr uint32_t kBBytes = kBRows * kBK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kNSub * kBBytes;   // 32 / 48 KB
  static constexpr int kStages = (kStageBytes == 32768) ? 6 : 4;
  static constexpr uint32_t kTxBytes = kPair ? 2 * kStageBytes : kStageBytes;
  static constexpr int kTileN = kBN * kNSub;                           // database rows per CTA tile
  static constexpr int kBufs = kNSub == 1 ? 2 : 1;                     // TMEM accumulator buffers
};

struct TcSmemTail {
  float rn[2][2 * kBN];
  unsigned long long prune_keys[4][kCandCap];
  float prune_thr[4];
  uint64_t full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};
template <bool kPair, int kNSub>
constexpr size_t tc_smem_bytes() {
  return 1024 + size_t(TcCfg<kPair, kNSub>::kStages) * TcCfg<kPair, kNSub>::kStageBytes + sizeof(TcSmemTail);
}

__device__ __forceinline__ uint32_t f2ord(float s) {      // order-preserving map float -> uint
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long cand_key(float s, uint32_t idx) {
  return (static_cast<unsigned long long>(f2ord(s)) << 32) | idx;
}

// Warp-cooperative prune of the candidate list of lane `L`'s row to its kKeep best entries (sorted ascending).
__device__ __forceinline__ void prune_row(int L, int lane, float* __restrict__ cs, uint32_t* __restrict__ ci,
                                          size_t base, int& cnt, float& thr, unsigned long long* keys, float* thr_slot,
                                          unsigned int* q_thr_of_lane) {
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
    // publish: every range of this query may now reject anything that scores >= the 64th best seen here
    atomicMin(q_thr_of_lane, f2ord(thr));
  }
  __syncwarp();
}

// Cheap prune: instead of ranking all ~128 entries (O(n^2/32) compares), sort a 32-entry sample with a shuffle network,
// take the sample quantile that should leave ~72 entries as the pivot, count, and compact everything <= pivot.  Any
// pivot is a VALID threshold (whatever is dropped or later rejected scores >= it, which is all the exactness
// certificate needs); the count only has to stay comfortably above k, otherwise the exact prune runs instead.
__device__ __forceinline__ void prune_row_fast(int L, int lane, float* __restrict__ cs, uint32_t* __restrict__ ci,
                                               size_t base, int& cnt, float& thr, unsigned long long* keys,
                                               float* thr_slot, unsigned int* q_thr_of_lane) {
  const int n = __shfl_sync(0xffffffffu, cnt, L);
  const unsigned long long b64 = static_cast<unsigned long long>(base);
  const size_t bL = static_cast<size_t>(__shfl_sync(0xffffffffu, b64, L));
  __syncwarp();
  float ms[kCandCap / 32];
  uint32_t mi[kCandCap / 32];
#pragma unroll
  for (int u = 0; u < kCandCap / 32; ++u) {
    const int e = lane + 32 * u;
    ms[u] = INFINITY;
    mi[u] = 0;
    if (e < n) {
      ms[u] = __ldcg(cs + bL + e);
      mi[u] = __ldcg(ci + bL + e);
    }
  }
  // bitonic sort of the 32 samples ms[0] (entries 0..31 exist: n > kCandCap - 32 >= 32)
  float v = ms[0];
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = ((lane & k) == 0);
      const bool lower = ((lane & j) == 0);
      v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
    }
  }
  int t = (32 * 72 + n / 2) / n;                    // sample rank that leaves ~72 of n
  t = t < 8 ? 8 : (t > 31 ? 31 : t);
  const float pivot = __shfl_sync(0xffffffffu, v, t);
  int c = 0;
#pragma unroll
  for (int u = 0; u < kCandCap / 32; ++u) c += __popc(__ballot_sync(0xffffffffu, ms[u] <= pivot));
  if (c < 48 || c > kCandCap - 40) {                // unlucky sample: exact prune (warp-uniform branch)
    prune_row(L, lane, cs, ci, base, cnt, thr, keys, thr_slot, q_thr_of_lane);
    return;
  }
  int run = 0;
#pragma unroll
  for (int u = 0; u < kCandCap / 32; ++u) {
    const bool keep = ms[u] <= pivot;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int o = run + __popc(bal & ((1u << lane) - 1u));
      cs[bL + o] = ms[u];
      ci[bL + o] = mi[u];
    }
    run += __popc(bal);
  }
  __syncwarp();
  if (lane == L) {
    cnt = c;
    thr = pivot;
    atomicMin(q_thr_of_lane, f2ord(pivot));
  }
  __syncwarp();
}

// Work item -> (database range, query unit).  Items are ordered group-major: a group of `group_m` query units runs
// against every range before the next group starts, so the CTAs that are resident at the same time share a small set
// of query blocks (kept hot in L2) and each database range is streamed from HBM once per group, by CTAs in lockstep.
__device__ __forceinline__ void decode_item(int item, int m_units, int NR, int group_m, int& range, int& mu) {
  const int per_group = group_m * NR;
  const int groups = (m_units + group_m - 1) / group_m;
  int g = item / per_group;
  if (g > groups - 1) g = groups - 1;
  const int rem = item - g * per_group;
  const int gsz = min(group_m, m_units - g * group_m);
  range = rem / gsz;
  mu = g * group_m + (rem - range * gsz);
}

template <bool kPair, int kNSub>
__global__ void __launch_bounds__(kTcThreads, 1) knn_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB, TcArgs a) {
  using Cfg = TcCfg<kPair, kNSub>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kStageBytes = Cfg::kStageBytes;
  constexpr int kTileN = Cfg::kTileN;
  constexpr uint32_t kBufs = Cfg::kBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(smem + size_t(kStages) * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;       // 0 = leader of the pair
  const int worker = kPair ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int num_workers = kPair ? int(gridDim.x >> 1) : int(gridDim.x);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tail->tmem_full[b], 1);
      mbar_init(&tail->tmem_empty[b], kPair ? 8 : 4);     // one arrival per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_2sm(&tail->tmem_base, kTmemCols); else tmem_alloc(&tail->tmem_base, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();       // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  const int num_k = a.Dp / kBK;
  const int m_units = kPair ? (a.num_m_blocks + 1) / 2 : a.num_m_blocks;   // query blocks (pairs of blocks) per range
  const int num_items = m_units * a.NR;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // pacing (see TcArgs::sync_ctr): only the leader CTA of a pair takes part, its peer follows through the ring
      const bool pace = a.sync_total > 0 && crank == 0;
      const int subs = a.sync_subs;
      int round = 0;
      int sp_next = 0;                                   // first sync point this worker has not bumped yet
      for (int item = worker; item < num_items; item += num_workers, ++round) {
        int range, mu;
        decode_item(item, m_units, a.NR, a.group_m, range, mu);
        const int mb = kPair ? 2 * mu + int(crank) : mu;
        const int t0 = range * a.tiles_per_range, t1 = min(a.num_n_tiles, t0 + a.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          for (int sub = 0; sub < subs; ++sub) {
            const int kc0 = sub * num_k / subs, kc1 = (sub + 1) * num_k / subs;
            const int sp = (round * a.tiles_per_range + (t - t0)) * subs + sub;
            if (pace) {
              while (sp_next < sp) atomicAdd(a.sync_ctr + sp_next++, 1u);      // points of a short range
              if (sp >= a.sync_window) {
                const volatile unsigned int* c = a.sync_ctr + (sp - a.sync_window);
                if (*c < unsigned(num_workers)) {
                  const long long w0 = clock64();
                  while (*c < unsigned(num_workers) && clock64() - w0 < 400000) __nanosleep(256);
                }
              }
            }
            for (int kc = kc0; kc < kc1; ++kc) {
              mbar_wait(&tail->empty[stage], phase ^ 1);
              uint8_t* sa = smem + size_t(stage) * kStageBytes;
              if (kPair) {
                // both CTAs report their bytes to the LEADER's barrier; only the leader arms it (with both shares)
                const uint32_t lead_bar = mapa_u32(smem_u32(&tail->full[stage]), 0);
                if (crank == 0) mbar_arrive_expect_tx(&tail->full[stage], Cfg::kTxBytes);
                tma_load_2d_2sm(sa, &tmA, lead_bar, kc * kBK, mb * kBM);
#pragma unroll
                for (int sub2 = 0; sub2 < kNSub; ++sub2)
                  tma_load_2d_2sm(sa + kABytes + sub2 * Cfg::kBBytes, &tmB, lead_bar, kc * kBK,
                                  t * kTileN + sub2 * kBN + int(crank) * int(Cfg::kBRows));
              } else {
                mbar_arrive_expect_tx(&tail->full[stage], Cfg::kTxBytes);
                tma_load_2d(sa, &tmA, &tail->full[stage], kc * kBK, mb * kBM);
#pragma unroll
                for (int sub2 = 0; sub2 < kNSub; ++sub2)
                  tma_load_2d(sa + kABytes + sub2 * Cfg::kBBytes, &tmB, &tail->full[stage], kc * kBK,
                              t * kTileN + sub2 * kBN);
              }
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (pace) { atomicAdd(a.sync_ctr + sp, 1u); sp_next = sp + 1; }
          }
        }
      }
      // this worker is done: it counts as "passed" for every remaining point, so nobody waits for it
      if (pace) while (sp_next < a.sync_total) atomicAdd(a.sync_ctr + sp_next++, 1u);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of a pair only) =====================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtF16, kPair ? 2 * kBM : kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_count = 0;
      for (int item = worker; item < num_items; item += num_workers) {
        int range, mu;
        decode_item(item, m_units, a.NR, a.group_m, range, mu);
        const int t0 = range * a.tiles_per_range, t1 = min(a.num_n_tiles, t0 + a.tiles_per_range);
        for (int t = t0; t < t1; ++t, ++tile_count) {
          const uint32_t buf = tile_count % kBufs, use = tile_count / kBufs;
          // kNSub == 1: two accumulators, tile i+1 goes to the one the epilogue drained two tiles ago.
          // kNSub == 2: ONE 512-column accumulator whose halves are released separately (tmem_empty[0] / [1]): the
          // epilogue drains the first half, the MMAs of the next tile start on it while the second half is drained
          if (kNSub == 1) {
            mbar_wait(&tail->tmem_empty[buf], (use & 1) ^ 1);     // epilogue(s) drained this accumulator
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + buf * kBN;
          for (int kc = 0; kc < num_k; ++kc) {
            mbar_wait(&tail->full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + size_t(stage) * kStageBytes);
            const uint64_t da = smem_desc_sw128(sa);
#pragma unroll
            for (int sub = 0; sub < kNSub; ++sub) {
              if (kNSub == 2 && kc == 0) {
                mbar_wait(&tail->tmem_empty[sub], (use & 1) ^ 1);  // this half of the accumulator is drained
                tc_fence_after();
              }
              const uint64_t db = smem_desc_sw128(sa + kABytes + sub * Cfg::kBBytes);
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k) {
                // advance 16 fp16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                if (kPair) mma_f16_ss_2sm(d_tmem + sub * kBN, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
                else mma_f16_ss(d_tmem + sub * kBN, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
              }
            }
            // frees the smem slot (in both CTAs) when these MMAs retire
            if (kPair) mma_commit_2sm(&tail->empty[stage], 0b11); else mma_commit(&tail->empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (kPair) mma_commit_2sm(&tail->tmem_full[buf], 0b11); else mma_commit(&tail->tmem_full[buf]);
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
    for (int item = worker; item < num_items; item += num_workers) {
      int range, mu;
      decode_item(item, m_units, a.NR, a.group_m, range, mu);
      const int mb = kPair ? 2 * mu + int(crank) : mu;
      const int t0 = range * a.tiles_per_range, t1 = min(a.num_n_tiles, t0 + a.tiles_per_range);
      const int q = mb * kBM + lq * 32 + lane;
      const bool valid = q < a.Q;
      const float qmul = valid ? a.qmul[q] : 0.0f;
      const size_t base = (size_t(valid ? q : 0) * a.NR + range) * kCandCap;
      unsigned int* my_thr = a.q_thr + (valid ? q : 0);
      const bool collect = a.collect != 0;
      float thr = collect ? (valid ? __ldg(a.fixed_thr + q) : -INFINITY) : INFINITY;
      int cnt = 0;
      for (int t = t0; t < t1; ++t, ++tile_count) {
        const uint32_t buf = tile_count % kBufs, use = tile_count / kBufs;
        const uint32_t rbuf = tile_count & 1;      // the |r|^2 staging is always double-buffered
        const int n0 = t * kTileN;
        // stage the |r|^2 of this tile's rows (rows beyond the shard score +inf)
#pragma unroll
        for (int u = 0; u < kTileN / 128; ++u) {
          const int r = n0 + u * 128 + et;
          tail->rn[rbuf][u * 128 + et] = r < a.R ? __ldg(a.rn + r) : INFINITY;
        }
        // thresholds published by other ranges of the same query (other CTAs) tighten this row's filter too
        if (valid && !collect) {
          const unsigned int pub = __ldcg(my_thr);
          if (pub != 0xffffffffu) thr = fminf(thr, ord2f(pub));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tail->tmem_full[buf], use & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(lq * 32) << 16) + buf * kBN;
#pragma unroll 1
        for (int c = 0; c < kTileN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          // scores of this row against 32 database rows; the common case (nothing beats the threshold) is branch-free
          float sc[32];
          {
            const float4* rn4 = reinterpret_cast<const float4*>(&tail->rn[rbuf][c * 32]);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 r = rn4[j4];
              sc[4 * j4 + 0] = fmaf(__uint_as_float(v[4 * j4 + 0]), qmul, r.x);
              sc[4 * j4 + 1] = fmaf(__uint_as_float(v[4 * j4 + 1]), qmul, r.y);
              sc[4 * j4 + 2] = fmaf(__uint_as_float(v[4 * j4 + 2]), qmul, r.z);
              sc[4 * j4 + 3] = fmaf(__uint_as_float(v[4 * j4 + 3]), qmul, r.w);
            }
          }
          if (a.dbg_scores != nullptr && valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c * 32 + j;
              if (n < a.R) a.dbg_scores[size_t(q) * a.R + n] = sc[j];
            }
          }
          uint32_t hit = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) hit |= (sc[j] < thr) ? (1u << j) : 0u;
          if (!valid) hit = 0;
          if (hit != 0) {
            if (collect) {
              // stage 2: one list per query shared by all ranges; survivors are rare (rows within the rounding bound
              // of the k-th neighbour), so one atomic per survivor is cheap
              while (hit) {
                const int j = __ffs(hit) - 1;
                hit &= hit - 1;
                const int pos = atomicAdd(a.coll_cnt + q, 1);
                if (pos < a.coll_cap) a.coll_idx[size_t(q) * a.coll_cap + pos] = uint32_t(n0 + c * 32 + j);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if ((hit >> j) & 1u) {
                  a.cand_s[base + cnt] = sc[j];
                  a.cand_i[base + cnt] = uint32_t(n0 + c * 32 + j);
                  ++cnt;
                }
              }
            }
          }
          // keep at least 32 free slots before the next chunk
          unsigned need = __ballot_sync(0xffffffffu, cnt > kCandCap - 32);
          while (need) {
            const int L = __ffs(need) - 1;
            prune_row_fast(L, lane, a.cand_s, a.cand_i, base, cnt, thr, keys, thr_slot, my_thr);
            need &= need - 1;
          }
          if (kNSub == 2 && (c & (kBN / 32 - 1)) == kBN / 32 - 1) {
            // a 256-column half of the accumulator has been read by this warp: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              const int half = c / (kBN / 32);
              if (kPair && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tail->tmem_empty[half]), 0));
              else mbar_arrive(&tail->tmem_empty[half]);
            }
          }
        }
        if (kNSub == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            // the accumulator of BOTH CTAs must be drained before the leader's MMA warp may overwrite it
            if (kPair && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tail->tmem_empty[buf]), 0));
            else mbar_arrive(&tail->tmem_empty[buf]);
          }
        }
      }
      // end of the item: publish the list length (<= kCandCap entries, unsorted; knn_cand_merge_kernel filters them by
      // the final published threshold and sorts what is left)
      if (valid && !collect) a.cand_cnt[size_t(q) * a.NR + range] = cnt;
      if (a.group_done != nullptr) {
        __threadfence();                     // this lane's list entries and count are visible device-wide ...
        __syncwarp();
        if (lane == 0) {                     // ... before the warp's arrival is
          const int groups = (m_units + a.group_m - 1) / a.group_m;
          int g = item / (a.group_m * a.NR);
          if (g > groups - 1) g = groups - 1;
          atomicAdd(a.group_done + g, 1u);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();       // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_atom32) {
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
                  swizzle_atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    set_last_error(msg, cudaErrorInvalidValue);
    return SCL_ERR_CUDA;
  }
  return SCL_OK;
}

int make_tmap_3d(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t batch, uint64_t row_pitch_bytes, uint64_t batch_pitch_bytes, uint32_t box_inner,
                 uint32_t box_outer, int swizzle_atom32) {
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
  cuuint64_t dims[3] = {inner, outer, batch};
  cuuint64_t strides[2] = {row_pitch_bytes, batch_pitch_bytes};
  cuuint32_t box[3] = {box_inner, box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                      : (swizzle_atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", int(r));
    set_last_error(msg, cudaErrorInvalidValue);
    return SCL_ERR_CUDA;
  }
  return SCL_OK;
}

static int tc_variant() {
  // 1 = single-CTA 128x256 tiles, 2 = CTA pairs 256x256 with double-buffered accumulators (default: the epilogue
  // of tile i overlaps the MMAs of tile i+1), 3 = CTA pairs 256x512 (a quarter less L2->SM traffic; ONE accumulator whose
  // halves are released separately, so the next tile's MMAs start on the first half while the second is drained).
  // Measured on a 125 000-row shard, 10 000 queries, one launch (round 2): variant 2 8.56 ms, variant 1 9.31 ms,
  // variant 3 10.9 ms before and 10.8 ms after the separate release of the halves, 11.8 ms with software-pipelined
  // tcgen05.ld in the epilogue (219-236 registers; reverted).  ncu of variant 3: L2->SM 62 GB (-25 %) at 6.2 TB/s,
  // tensor pipe 50 %: it is neither L2- nor tensor-bound, its MMA warp waits for the 512-column drain.
  const int v = knob(KNOB_KNN_TC_VARIANT);
  return (v >= 1 && v <= 3) ? v : 2;
}
static int tc_tile_n(int variant) { return variant == 3 ? 2 * kBN : kBN; }

template <bool kPair, int kNSub>
static int tc_launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, cudaStream_t stream) {
  auto kern = knn_tc_kernel<kPair, kNSub>;
  static SmemAttrCache configured;                     // per device
  int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), tc_smem_bytes<kPair, kNSub>(), &configured);
  if (rc_attr) return rc_attr;
  const int sms = num_sms();
  const int m_units = kPair ? (a.num_m_blocks + 1) / 2 : a.num_m_blocks;
  const int items = m_units * a.NR;
  const int workers_max = kPair ? sms / 2 : sms;
  const int workers = items < workers_max ? items : workers_max;
  TcArgs args = a;
  {
    // pacing of the producers: `subs` sync points per tile, window in sync points (SCL_KNN_SYNC=0 switches it off)
    const int num_k = a.Dp / kBK;
    int subs = knob_or(KNOB_KNN_SYNC_SUBS, 4);
    if (subs < 1) subs = 1;
    if (subs > num_k) subs = num_k;
    int window = knob_or(KNOB_KNN_SYNC_WINDOW, subs);
    if (window < 1) window = 1;
    const long long rounds = (items + workers - 1) / workers;
    const long long total = rounds * a.tiles_per_range * subs;
    const bool on = knob_or(KNOB_KNN_SYNC, 1) != 0 && a.sync_ctr != nullptr && workers > 1 && total <= kSyncMax;
    args.sync_subs = subs;
    args.sync_window = window;
    args.sync_total = on ? int(total) : 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(kPair ? 2 * workers : workers));
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = tc_smem_bytes<kPair, kNSub>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SCL_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, args));
  return SCL_OK;
}

int knn_tc_launch(const TcArgs& a, const void* qh, const void* dbh, cudaStream_t stream) {
  const int variant = tc_variant();
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, qh, uint64_t(a.Dp), uint64_t(a.Q),
                        uint64_t(a.Dp) * 2, kBK, kBM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dbh, uint64_t(a.Dp), uint64_t(a.R), uint64_t(a.Dp) * 2,
                    kBK, variant == 1 ? kBN : kBN / 2);
  if (rc) return rc;
  if (variant == 1) return tc_launch_variant<false, 1>(tmA, tmB, a, stream);
  if (variant == 2) return tc_launch_variant<true, 1>(tmA, tmB, a, stream);
  return tc_launch_variant<true, 2>(tmA, tmB, a, stream);
}

void knn_tc_tiling(int Q, int64_t R, int Dp, int* num_m_blocks, int* num_n_tiles, int* NR, int* tiles_per_range, int* group_m) {
  const int mb = (Q + kBM - 1) / kBM;
  const int variant = tc_variant();
  const int tile_n = tc_tile_n(variant);
  const int nt = int((R + tile_n - 1) / tile_n);
  const bool pair = variant >= 2;
  const int sms = pair ? num_sms() / 2 : num_sms();      // workers: CTA pairs or single CTAs
  const int mu = pair ? (mb + 1) / 2 : mb;               // work units along the query axis
  // choose the number of database ranges so that (query blocks x ranges) fills whole waves of SMs
  int best = 1;
  double best_eff = -1.0;
  const int max_nr = nt < 64 ? nt : 64;
  for (int nr = 1; nr <= max_nr; ++nr) {
    const long long items = 1ll * mu * nr;
    const long long waves = (items + sms - 1) / sms;
    const int tpr = (nt + nr - 1) / nr;
    // time ~ waves * tiles_per_range; efficiency relative to perfect balance
    const double eff = double(1ll * mu * nt) / (double(waves) * sms * tpr);
    if (eff > best_eff + 0.005) { best_eff = eff; best = nr; }
  }
  const int forced_nr = knob(KNOB_KNN_RANGES);
  if (forced_nr >= 1 && forced_nr <= max_nr) best = forced_nr;
  *num_m_blocks = mb;
  *num_n_tiles = nt;
  *NR = best;
  *tiles_per_range = (nt + best - 1) / best;
  // query-axis work units per group: ~40 MB of fp16 query blocks (20 CTA-pair units at D = 4096) stay L2-resident next
  // to the database streams, and each range is streamed from HBM only ceil(units / group) times (measured on B200:
  // group 20 -> 1267 TFLOP/s, 8 -> 1160, 40 -> 1119)
  const long long unit_bytes = (pair ? 2ll : 1ll) * kBM * Dp * 2;
  int gm = int((40ll << 20) / (unit_bytes > 0 ? unit_bytes : 1));
  if (gm < 2) gm = 2;
  if (knob(KNOB_KNN_GROUP_M) >= 1) gm = knob(KNOB_KNN_GROUP_M);
  if (gm > mu) gm = mu;
  gm = (mu + (mu + gm - 1) / gm - 1) / ((mu + gm - 1) / gm);      // equal-sized groups
  *group_m = gm;
}

}  // namespace scl
