// tuple_common.cuh -- pieces shared by the tuple-mode loss kernels (wms_tuple.cu, tuple_losses.cu).
//
// Every tuple-mode loss in /root/reference/model/losses.py has the same data-flow shape on a tuple of S
// descriptors e_0..e_{S-1}:  (1) a handful of pairwise reductions over the descriptor dimension,
// (2) O(S^2) scalar logic, (3) a gradient whose every row is a linear combination of the tuple's rows,
// d loss / d E = M * E with an S x S coefficient matrix M.  One thread-block cluster owns one tuple; each CTA
// owns a column slice that stays in shared memory between (1) and (3), so HBM sees E once and dE once.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace scl {

namespace cg = cooperative_groups;

constexpr int kTupThreads = 256;
constexpr int kTupWarps = kTupThreads / 32;
constexpr int kMaxCluster = 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__host__ __device__ constexpr size_t al4(size_t n) { return (n + 3) / 4 * 4; }   // float counts -> 16-byte multiples

// Row-count bucket: SG padded rows, processed as two halves of HR rows in the backward.
template <int SG_>
struct TupDims {
  static constexpr int SG = SG_;
  static constexpr int HR = (SG + 1) / 2;
  static constexpr int HRP = (HR + 3) / 4 * 4;
  static constexpr int MT = SG * 2 * HRP;   // floats of the transposed coefficient matrix
};

// Bring rows [0,S) x columns [col0, col0+Dc) of one tuple into Es (row pitch Dc+4 floats).
__device__ __forceinline__ void tup_load_chunk(float* Es, const float* src, int S, int D, int Dc) {
  const int ncols4 = Dc >> 2, pitch = Dc + 4;
  const int total = S * ncols4;
  for (int i = threadIdx.x; i < total; i += kTupThreads) {
    int r = i / ncols4, c4 = i - r * ncols4;
    cp_async16(Es + r * pitch + 4 * c4, src + size_t(r) * D + 4 * c4);
  }
  cp_async_wait_all();
  __syncthreads();
}

// Mt[j][h*HRP + r] = M[h*HR + r][j] * scale   (zero-padded).  Call with all threads; Mraw is [S][ldm].
template <int SG>
__device__ __forceinline__ void tup_store_Mt(float* Mt, const float* Mraw, int ldm, int S) {
  using Dm = TupDims<SG>;
  for (int k = threadIdx.x; k < Dm::MT; k += kTupThreads) Mt[k] = 0.0f;
  __syncthreads();
  for (int k = threadIdx.x; k < S * S; k += kTupThreads) {
    int i = k / S, j = k - i * S;
    int h = i / Dm::HR, r = i - h * Dm::HR;
    Mt[j * (2 * Dm::HRP) + h * Dm::HRP + r] = Mraw[i * ldm + j];
  }
  __syncthreads();
}

// dE[i][cols] = scale * sum_j M[i][j] * E[j][cols] for the chunk resident in Es.
template <int SG>
__device__ __forceinline__ void tup_bwd_chunk(const float* Es, const float* Mt, float* dE_chunk, int S, int D, int Dc,
                                              float scale) {
  using Dm = TupDims<SG>;
  constexpr int HR = Dm::HR, HRP = Dm::HRP;
  const int ncols4 = Dc >> 2, pitch = Dc + 4;
  for (int item = threadIdx.x; item < 2 * ncols4; item += kTupThreads) {
    const int h = item / ncols4, c4 = item - h * ncols4;
    float4 o[HR];
#pragma unroll
    for (int r = 0; r < HR; ++r) o[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < S; ++j) {
      const float4 e = *reinterpret_cast<const float4*>(Es + j * pitch + 4 * c4);
      const float4* mrow = reinterpret_cast<const float4*>(Mt + j * (2 * HRP) + h * HRP);
#pragma unroll
      for (int r4 = 0; r4 < HRP / 4; ++r4) {
        const float4 m = mrow[r4];
        const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r4 * 4 + u;
          if (r < HR) {
            o[r].x = fmaf(mv[u], e.x, o[r].x);
            o[r].y = fmaf(mv[u], e.y, o[r].y);
            o[r].z = fmaf(mv[u], e.z, o[r].z);
            o[r].w = fmaf(mv[u], e.w, o[r].w);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      const int i = h * HR + r;
      if (i < S) {
        float4 v = make_float4(o[r].x * scale, o[r].y * scale, o[r].z * scale, o[r].w * scale);
        stg_stream(reinterpret_cast<float4*>(dE_chunk + size_t(i) * D + 4 * c4), v);
      }
    }
  }
}

// Deterministic batch reduction, called by one full warp per tuple: lane 0 deposits the tuple's value; the
// warp of the last tuple to arrive sums all of them in a fixed lane-strided order and writes the mean.
// `ws` = {u32 counter (zeroed by the host before launch), pad[3], float stage[T]}.
__device__ __forceinline__ void tup_finish_loss(unsigned int* ws, int t, int T, float v, float* loss_out, int lane) {
  float* stage = reinterpret_cast<float*>(ws + 4);
  unsigned prev = 0;
  if (lane == 0) {
    stage[t] = v;
    __threadfence();
    prev = atomicAdd(ws, 1u);
  }
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev == unsigned(T) - 1u) {
    __threadfence();
    float tot = 0.0f;
    for (int k = lane; k < T; k += 32) tot += __ldcg(stage + k);
    tot = warp_sum(tot);
    if (lane == 0) loss_out[0] = tot / float(T);
  }
}

// The same for persistent CTAs that own many tuples: the per-tuple call only deposits the value (a plain store -- no
// fence, no atomic round trip on the tuple's critical path); after its last tuple the CTA's depositing warp publishes
// all of them with ONE fence + atomic and the last CTA to arrive reduces in the same fixed order.
__device__ __forceinline__ void tup_deposit_loss(unsigned int* ws, int t, float v, int lane) {
  if (lane == 0) reinterpret_cast<float*>(ws + 4)[t] = v;
}
__device__ __forceinline__ void tup_publish_losses(unsigned int* ws, int n_mine, int T, float* loss_out, int lane) {
  float* stage = reinterpret_cast<float*>(ws + 4);
  unsigned prev = 0;
  if (lane == 0) {
    __threadfence();
    prev = atomicAdd(ws, unsigned(n_mine));
  }
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev + unsigned(n_mine) == unsigned(T)) {
    __threadfence();
    float tot = 0.0f;
    for (int k = lane; k < T; k += 32) tot += __ldcg(stage + k);
    tot = warp_sum(tot);
    if (lane == 0) loss_out[0] = tot / float(T);
  }
}

// Host-side plan shared by the tuple kernels.
struct TupPlan {
  int sg;        // 25, 30 or 35
  int cluster;   // 1,2,4,8
  int Ds, Dc;
  size_t smem;
};

}  // namespace scl
