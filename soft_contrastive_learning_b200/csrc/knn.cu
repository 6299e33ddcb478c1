// knn.cu -- R1/R2/R3: exact brute-force kNN over a database shard, shard merge, geo bookkeeping, recall.
//
// Replaces KDTree(ref_f).query(query_f, k=N, return_distance=True, sort_results=True)
// (/root/reference/evaluation/top-n.py:103-106; train/train.py:1181-1182, :451) and the bookkeeping of
// top-n.py:69,110-117 / roc.py:213-216 / train.py:368-375.
//
// Two device paths, both returning EXACT float64 results ordered by (distance, index):
//   exact scan   d^2(q,r) = sum (q_i - r_i)^2 accumulated in float64 for every row, radix-select of the k smallest.
//   tensor pass  (knn_tc.cu) fp16 scores on tcgen05 keep k'=64 candidates per (query, database range);
//                here: merge the ranges, rescore the 64 best candidates exactly (float64), order them and
//                CERTIFY each query: with T the 64th smallest fp16 score and eps a rigorous bound on
//                |fp16 score - exact score|, every row that is not a candidate has d^2 >= T + |q|^2 - eps, so
//                when the k-th exact candidate distance is below that, the exact top-k is inside the candidate set.
//                Queries that fail the certificate are recomputed by the exact scan.  No approximation is returned.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "knn_internal.cuh"

namespace scl {

// ---------------------------------------------------------------------------------------------
// index build
// ---------------------------------------------------------------------------------------------
static inline size_t shadow_norm_off() { return sizeof(ShadowHeader); }
static inline size_t shadow_data_off(int64_t R) { return sizeof(ShadowHeader) + round_up(size_t(R) * sizeof(float), 1024); }
static inline int pad64(int D) { return (D + 63) / 64 * 64; }

__global__ void knn_header_init_kernel(ShadowHeader* h, long long R, int D, int Dp) {
  h->R = R; h->D = D; h->Dp = Dp;
  h->maxabs_bits = 0u; h->scale_exp = 0; h->rmax2_bits = 0ull; h->built = 0;
}

// pass 1: exact squared norms (float64 accumulate), max |x|, max norm.  One warp per row.
__global__ void __launch_bounds__(256) knn_build_stats_kernel(const float* __restrict__ db, long long R, int D,
                                                              float* __restrict__ rn, ShadowHeader* h) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float mx = 0.0f;
  double rmax = 0.0;
  for (long long r = warp; r < R; r += nw) {
    const float4* row = reinterpret_cast<const float4*>(db + size_t(r) * D);
    double acc = 0.0;
    for (int c = lane; c < (D >> 2); c += 32) {
      const float4 v = ldg_stream(row + c);
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      acc = fma(double(v.x), double(v.x), acc);
      acc = fma(double(v.y), double(v.y), acc);
      acc = fma(double(v.z), double(v.z), acc);
      acc = fma(double(v.w), double(v.w), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) rn[r] = float(acc);
    rmax = fmax(rmax, acc);
  }
  mx = warp_max(mx);
  if (lane == 0) {
    atomicMax(&h->maxabs_bits, __float_as_uint(mx));
    atomicMax(&h->rmax2_bits, (unsigned long long)__double_as_longlong(rmax));
  }
}

// power-of-two scale that puts max|x| in [2^13, 2^14): fp16 keeps 11 significant bits for everything within
// 2^-27 of the largest magnitude and cannot overflow
__device__ __forceinline__ int scale_exp_for(float maxabs) {
  if (!(maxabs > 0.0f) || isinf(maxabs)) return 0;
  return 13 - ilogbf(maxabs);
}

__global__ void knn_build_scale_kernel(ShadowHeader* h) {
  h->scale_exp = scale_exp_for(__uint_as_float(h->maxabs_bits));
  h->built = 1;
}

// pass 2: fp16 shadow rows (pitch Dp, zero padded)
__global__ void __launch_bounds__(256) knn_build_convert_kernel(const float* __restrict__ db, long long R, int D, int Dp,
                                                                const ShadowHeader* __restrict__ h,
                                                                __half* __restrict__ out) {
  const float sc = ldexpf(1.0f, h->scale_exp);
  const long long total4 = R * (long long)(Dp >> 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (Dp >> 2);
    const int c = int(i - r * (Dp >> 2)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) v = ldg_stream(reinterpret_cast<const float4*>(db + size_t(r) * D + c));
    __half2 lo = __floats2half2_rn(v.x * sc, v.y * sc), hi = __floats2half2_rn(v.z * sc, v.w * sc);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + size_t(r) * Dp + c) = pk;
  }
}

// ---------------------------------------------------------------------------------------------
// query preparation: fp16 copy with a per-query power-of-two scale, |q|^2 in float64
// ---------------------------------------------------------------------------------------------
// One CTA per query, the row held in registers between the two steps (max |q| / sum q^2, then scale + convert): one
// pass over HBM with 16-byte loads and 8-byte stores for D <= 4096 (kPrepVec float4 per thread); the tail of a longer
// row is read a second time (L2).
constexpr int kPrepVec = 4;                           // float4 per thread in registers: D <= 4096 in one pass
__global__ void __launch_bounds__(256) knn_query_prep_kernel(const float* __restrict__ q, int Q, int D, int Dp,
                                                             const ShadowHeader* __restrict__ h, __half* __restrict__ qh,
                                                             float* __restrict__ qmul, double* __restrict__ qn2,
                                                             int* __restrict__ qexp) {
  __shared__ float s_mx[8];
  __shared__ double s_acc[8];
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4* row4 = reinterpret_cast<const float4*>(q + size_t(qi) * D);
  const int nv = D >> 2;                              // D % 4 == 0 (checked by the entry points)
  float4 v[kPrepVec];
  float mx = 0.0f;
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < kPrepVec; ++u) {
    const int c = threadIdx.x + u * 256;
    v[u] = c < nv ? __ldg(row4 + c) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
#pragma unroll
  for (int u = 0; u < kPrepVec; ++u) {
    mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[u].x), fabsf(v[u].y))), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
    acc = fma(double(v[u].x), double(v[u].x), acc); acc = fma(double(v[u].y), double(v[u].y), acc);
    acc = fma(double(v[u].z), double(v[u].z), acc); acc = fma(double(v[u].w), double(v[u].w), acc);
  }
  for (int c = threadIdx.x + kPrepVec * 256; c < nv; c += 256) {       // rows longer than the register tile
    const float4 t = __ldg(row4 + c);
    mx = fmaxf(fmaxf(mx, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
    acc = fma(double(t.x), double(t.x), acc); acc = fma(double(t.y), double(t.y), acc);
    acc = fma(double(t.z), double(t.z), acc); acc = fma(double(t.w), double(t.w), acc);
  }
  mx = warp_max(mx);
  acc = warp_sum(acc);
  if (lane == 0) { s_mx[warp] = mx; s_acc[warp] = acc; }
  __syncthreads();
  mx = 0.0f; acc = 0.0;
  for (int w = 0; w < (blockDim.x >> 5); ++w) { mx = fmaxf(mx, s_mx[w]); acc += s_acc[w]; }
  const int e = scale_exp_for(mx);
  const float sc = ldexpf(1.0f, e);
  uint2* out = reinterpret_cast<uint2*>(qh + size_t(qi) * Dp);         // Dp % 64 == 0: 8-byte groups of four halves
  auto pack = [&](const float4& t) {
    const __half2 lo = __floats2half2_rn(t.x * sc, t.y * sc), hi = __floats2half2_rn(t.z * sc, t.w * sc);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  };
#pragma unroll
  for (int u = 0; u < kPrepVec; ++u) {
    const int c = threadIdx.x + u * 256;
    if (c < nv) out[c] = pack(v[u]);
  }
  for (int c = threadIdx.x + kPrepVec * 256; c < nv; c += 256) out[c] = pack(__ldg(row4 + c));
  for (int c = nv + threadIdx.x; c < (Dp >> 2); c += 256) out[c] = make_uint2(0u, 0u);      // zero padding up to Dp
  if (threadIdx.x == 0) {
    qmul[qi] = -2.0f * ldexpf(1.0f, -(e + h->scale_exp));
    qn2[qi] = acc;
    qexp[qi] = e;
  }
}

// ---------------------------------------------------------------------------------------------
// tensor pass, stage 2: merge the per-range candidate lists of one query, keep the kKeep best fp16 scores
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cand_key64(float s, uint32_t idx) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | idx;
}
__device__ __forceinline__ float key64_score(unsigned long long k) {
  uint32_t u = uint32_t(k >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}


template <typename K>
__device__ __forceinline__ void bitonic_sort_smem(K* keys, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        K a = keys[lo], b = keys[hi];
        if ((b < a) == up) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// rigorous bound on |fp16-pass score - exact score| for query q against ANY row of the shard
__device__ double knn_eps(double qnorm, int eq, const ShadowHeader* h) {
  const double u = ldexp(1.0, -11), a = ldexp(1.0, -25);
  const double rmax = sqrt(__longlong_as_double((long long)h->rmax2_bits));
  const double Qn = ldexp(qnorm, eq), Rn = ldexp(rmax, h->scale_exp);
  const double sD = sqrt(double(h->Dp));
  const double Qt = Qn * (1.0 + u) + a * sD, Rt = Rn * (1.0 + u) + a * sD;        // norms of the rounded vectors
  double e = u * (2.0 + u) * Qn * Rn + a * (1.0 + u) * sD * (Qn + Rn) + double(h->Dp) * a * a;   // input rounding
  e += (double(h->Dp) / 16.0 + 1.0) * ldexp(1.0, -21) * Qt * Rt;                // tensor-core fp32 accumulation
  const double e_dot = ldexp(e, -(eq + h->scale_exp));
  const double smax = rmax * rmax + 2.0 * qnorm * rmax + 2.0 * e_dot;
  return 2.0 * e_dot + ldexp(1.0, -23) * (rmax * rmax + smax);                    // + norm and fma rounding
}

// One CTA per query.  Three dependent trips to memory in total (counts, scores, indices of the selected): every thread
// first pulls its <= 32 list slots into registers with independent loads.  Only candidates at or below the final
// published threshold can be among the kKeep best of the union (the range that published it holds kKeep entries at or
// below it); of those survivors (a few hundred to a few thousand) the kKeep smallest are found with a block-wide
// 4 x 8-bit radix select over the register-resident scores -- no shared-memory copy of the lists, no large sort --
// and only that small set (plus score ties at the boundary) is fetched, sorted by (score, index) and cut at kKeep.
constexpr int kMergeCap = 256;                       // selected keys: kKeep + boundary ties; more ties => exact path (guard)
constexpr int kMergeSlots = 64 * kCandCap / 256;     // list slots per thread at the maximum of 64 ranges
__device__ __forceinline__ uint32_t score_key32(float s) {
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256, 3) knn_cand_merge_kernel(const float* __restrict__ cand_s,
                                                             const uint32_t* __restrict__ cand_i,
                                                             const int* __restrict__ cand_cnt,
                                                             const unsigned int* __restrict__ q_thr, int NR, int k,
                                                             const double* __restrict__ qn2, const int* __restrict__ qexp,
                                                             const ShadowHeader* __restrict__ h,
                                                             uint32_t* __restrict__ sel_idx, float* __restrict__ sel_T,
                                                             int* __restrict__ sel_n, float* __restrict__ sel_score,
                                                             float* __restrict__ bound_out) {
  // sel_score / bound_out != nullptr: first phase of the sharded two-phase query (scl_knn_query_begin): every one of
  // the kKeep best candidates is kept with its fp16 score (the cut is applied later, against the bound the ranks agree
  // on), and bound_out[q, j] = s_j + eps for the k best (ascending): each an upper bound, in score units, on the exact
  // distance of one actual row of this shard (+inf where the shard has fewer).
  const bool two_phase = sel_score != nullptr;
  __shared__ unsigned long long mkeys[kMergeCap];
  __shared__ int s_cnt[64];
  __shared__ int s_hist[256];
  __shared__ int s_need, s_total, s_surv, s_rank;
  __shared__ uint32_t s_prefix;
  __shared__ double s_cut;
  const int q = blockIdx.x;
  if (threadIdx.x < 64) s_cnt[threadIdx.x] = threadIdx.x < NR ? cand_cnt[size_t(q) * NR + threadIdx.x] : 0;   // NR <= 64
  const unsigned int pub = q_thr[q];
  if (threadIdx.x == 0) { s_need = 0; s_surv = 0; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int t = __reduce_add_sync(0xffffffffu, s_cnt[threadIdx.x] + s_cnt[threadIdx.x + 32]);
    if (threadIdx.x == 0) s_total = t;
  }
  const float t_pub = pub == 0xffffffffu ? INFINITY : key64_score((unsigned long long)pub << 32);
  const float* qs = cand_s + size_t(q) * NR * kCandCap;
  const uint32_t* qi = cand_i + size_t(q) * NR * kCandCap;
  uint32_t ku[kMergeSlots];                           // monotone keys of this thread's slots
  uint32_t live = 0;                                  // bit it: slot holds a survivor
  // the slot loops run over the groups of 8 slots per thread that the NR ranges of this launch reach (fully unrolled:
  // ku stays in registers); NR = 37 at an 8-way shard of config 4 -> 3 of the 4 groups
  const int nslots = (NR * kCandCap + 255) / 256;
#define SCL_FOR_SLOTS(body)                                                  \
  _Pragma("unroll") for (int qd = 0; qd < kMergeSlots / 8; ++qd) {           \
    if (qd * 8 < nslots) {                                                   \
      _Pragma("unroll") for (int j = 0; j < 8; ++j) { const int it = qd * 8 + j; body }  \
    }                                                                        \
  }
  SCL_FOR_SLOTS(
    const int slot = it * 256 + threadIdx.x; const int r = slot / kCandCap; const int e = slot % kCandCap;
    const float sc = (r < NR && e < s_cnt[r]) ? qs[slot] : NAN;          // NaN: never <= anything
    ku[it] = score_key32(sc);
    live |= (sc <= t_pub) ? (1u << it) : 0u;)
  if (live) atomicAdd(&s_surv, __popc(live));
  __syncthreads();
  const int kept = s_surv;                            // survivors of the threshold
  // radix select of the kKeep-th smallest key (only when there is something to cut)
  uint32_t tau = 0xffffffffu;                         // keys < tau are selected outright, keys == tau are boundary ties
  if (kept > kMergeCap) {
    if (threadIdx.x == 0) { s_rank = kKeep; s_prefix = 0; }
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
      s_hist[threadIdx.x] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      SCL_FOR_SLOTS(if (((live >> it) & 1u) && (shift == 24 || (ku[it] >> (shift + 8)) == prefix))
                      atomicAdd(&s_hist[(ku[it] >> shift) & 255u], 1);)
      __syncthreads();
      if (threadIdx.x < 32) {
        // bins 8*lane .. 8*lane+7 per lane, exclusive scan across lanes, then locate the bin holding rank s_rank
        int c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = s_hist[8 * threadIdx.x + j]; sum += c[j]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (int(threadIdx.x) >= o) incl += v;
        }
        int before = incl - sum;
        const int rank = s_rank;
        __syncwarp();                                 // every lane has read s_rank / s_prefix before one lane rewrites them
        if (before < rank && rank <= incl) {          // exactly one lane
          int j = 0;
          while (before + c[j] < rank) { before += c[j]; ++j; }
          s_rank = rank - before;
          s_prefix = (prefix << 8) | uint32_t(8 * threadIdx.x + j);
        }
      }
      __syncthreads();
    }
    tau = s_prefix;
  }
  // gather the selected keys (everything when the survivors fit)
  int mine = 0;
  SCL_FOR_SLOTS(mine += (((live >> it) & 1u) && ku[it] <= tau) ? 1 : 0;)
  int o = mine ? atomicAdd(&s_need, mine) : 0;
  SCL_FOR_SLOTS(if (((live >> it) & 1u) && ku[it] <= tau) {
    if (o < kMergeCap) mkeys[o] = (static_cast<unsigned long long>(ku[it]) << 32) | qi[it * 256 + threadIdx.x];
    ++o;
  })
#undef SCL_FOR_SLOTS
  __syncthreads();
  const int total = s_total;
  const bool overflow = s_need > kMergeCap;          // too many score ties at the boundary: handed to the exact path
  const int gathered = overflow ? kMergeCap : s_need;
  int n_pow2 = kKeep;
  while (n_pow2 < gathered) n_pow2 <<= 1;
  for (int i = gathered + threadIdx.x; i < n_pow2; i += blockDim.x) mkeys[i] = ~0ull;
  bitonic_sort_smem(mkeys, n_pow2);
  // Everything that is NOT among the selected entries scored >= T:
  //   entries left in the lists score >= the kKeep-th smallest of the union;
  //   anything a range rejected or pruned scored >= the threshold in force, which was >= the final published one.
  // Of the kKeep best, only candidates within 2*eps of the k-th best fp16 score can belong to the exact top-k
  // (a candidate c beyond that has exact_c >= s_c - eps > s_k + eps >= the exact score of each of the k best), so the
  // rest is not rescored and T is lowered to that cutoff.
  const float t_sel = kept >= kKeep ? key64_score(mkeys[kKeep - 1]) : INFINITY;
  float T = fminf(t_sel, t_pub);
  int n_sel = kept < kKeep ? kept : kKeep;
  if (two_phase) {
    if (threadIdx.x < kKeep)
      sel_score[size_t(q) * kKeep + threadIdx.x] = threadIdx.x < n_sel ? key64_score(mkeys[threadIdx.x]) : INFINITY;
    if (threadIdx.x == 0) s_cut = knn_eps(sqrt(qn2[q]), qexp[q], h);
    __syncthreads();
    if (threadIdx.x < k)                                // rounded up: a larger bound only keeps more candidates
      bound_out[size_t(q) * k + threadIdx.x] = (threadIdx.x < n_sel && !overflow)
          ? __double2float_ru(double(key64_score(mkeys[threadIdx.x])) + s_cut) : INFINITY;
  } else if (total > k && (long long)total < h->R) {
    if (threadIdx.x == 0) {                            // float64 evaluation of the bound: one thread, not 256
      s_need = 0;
      s_cut = double(key64_score(mkeys[k - 1])) + 2.0 * knn_eps(sqrt(qn2[q]), qexp[q], h);
    }
    __syncthreads();
    const double cut = s_cut;
    if (threadIdx.x < n_sel && double(key64_score(mkeys[threadIdx.x])) <= cut) atomicAdd(&s_need, 1);
    __syncthreads();
    if (s_need < n_sel) {
      n_sel = s_need;                                  // keys are sorted: the needed ones are a prefix
      T = fminf(T, __double2float_rd(cut));
    }
  }
  if (threadIdx.x < kKeep)
    sel_idx[size_t(q) * kKeep + threadIdx.x] = (threadIdx.x < n_sel) ? uint32_t(mkeys[threadIdx.x] & 0xffffffffu) : 0xffffffffu;
  if (threadIdx.x == 0) {
    sel_n[q] = overflow ? 0x7fffffff : total;
    sel_T[q] = overflow ? -INFINITY : T;
  }
}

// exact squared distances of the selected candidates, float64 accumulation of direct differences (what KDTree's leaf
// scan computes).  One CTA per query, one warp per candidate row, the warps striding over the list: the selected
// candidates are a PREFIX of the kKeep slots (sorted by score, cut from the top), so a warp stops at its first empty slot
// -- a two-phase (sharded) call rescoring ~k/G rows per query launches no work for the other slots.  Empty slots are not
// written (knn_finalize_kernel reads them as +inf from sel_idx).
__global__ void __launch_bounds__(256) knn_rescore_kernel(const float* __restrict__ db, const float* __restrict__ q,
                                                          int D, const uint32_t* __restrict__ sel_idx,
                                                          double* __restrict__ d2) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t qi = blockIdx.x;
  const float4* qr = reinterpret_cast<const float4*>(q + qi * D);
  const int nv = D >> 2;
  for (int s = warp; s < kKeep; s += 8) {
    const uint32_t r = sel_idx[qi * kKeep + s];
    if (r == 0xffffffffu) break;
    const float4* rr = reinterpret_cast<const float4*>(db + size_t(r) * D);
    double acc = 0.0;
    int c = lane;
    for (; c + 96 < nv; c += 128) {                     // four independent 16-byte loads in flight per lane
      const float4 b0 = ldg_stream(rr + c), b1 = ldg_stream(rr + c + 32), b2 = ldg_stream(rr + c + 64),
                   b3 = ldg_stream(rr + c + 96);
      const float4 a0 = __ldg(qr + c), a1 = __ldg(qr + c + 32), a2 = __ldg(qr + c + 64), a3 = __ldg(qr + c + 96);
      double d;
#define SCL_ACC4(a, b)                                              \
      d = double(a.x) - double(b.x); acc = fma(d, d, acc);          \
      d = double(a.y) - double(b.y); acc = fma(d, d, acc);          \
      d = double(a.z) - double(b.z); acc = fma(d, d, acc);          \
      d = double(a.w) - double(b.w); acc = fma(d, d, acc);
      SCL_ACC4(a0, b0) SCL_ACC4(a1, b1) SCL_ACC4(a2, b2) SCL_ACC4(a3, b3)
    }
    for (; c < nv; c += 32) {
      const float4 a0 = __ldg(qr + c);
      const float4 b0 = ldg_stream(rr + c);
      double d;
      SCL_ACC4(a0, b0)
    }
#undef SCL_ACC4
    acc = warp_sum(acc);
    if (lane == 0) d2[qi * kKeep + s] = acc;
  }
}

// order the kKeep rescored candidates of one query by (d^2, idx), emit the top-k, certify.  One warp per query.
__global__ void __launch_bounds__(256) knn_finalize_kernel(const uint32_t* __restrict__ sel_idx,
                                                           const double* __restrict__ d2, const float* __restrict__ sel_T,
                                                           const int* __restrict__ sel_n, const double* __restrict__ qn2,
                                                           const int* __restrict__ qexp, const ShadowHeader* __restrict__ h,
                                                           int Q, int k, long long idx_offset, int force_all, int q0,
                                                           double* __restrict__ out_d, long long* __restrict__ out_i,
                                                           double* __restrict__ kth_d2, int* __restrict__ flag_list,
                                                           int* __restrict__ stats, const int* __restrict__ cut_mode) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  double md[2];
  uint32_t mi[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    mi[u] = sel_idx[size_t(q) * kKeep + lane + 32 * u];
    md[u] = mi[u] != 0xffffffffu ? d2[size_t(q) * kKeep + lane + 32 * u] : INFINITY;      // empty slots are not rescored
  }
  int rank[2] = {0, 0};
  for (int src = 0; src < 32; ++src) {
#pragma unroll
    for (int su = 0; su < 2; ++su) {
      const double od = __shfl_sync(0xffffffffu, md[su], src);
      const uint32_t oi = __shfl_sync(0xffffffffu, mi[su], src);
#pragma unroll
      for (int u = 0; u < 2; ++u)           // empty slots (all idx 0xffffffff, d2 inf) are ordered by position
        rank[u] += (od < md[u] || (od == md[u] && (oi < mi[u] || (oi == mi[u] && src + 32 * su < lane + 32 * u)))) ? 1 : 0;
    }
  }
  double kth = INFINITY;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (rank[u] < k) {
      const bool real = mi[u] != 0xffffffffu;
      out_d[size_t(q) * k + rank[u]] = real ? sqrt(md[u]) : INFINITY;
      out_i[size_t(q) * k + rank[u]] = real ? (long long)mi[u] + idx_offset : -1ll;
    }
    if (rank[u] == k - 1) kth = md[u];
  }
  // broadcast the k-th exact distance
  for (int o = 16; o > 0; o >>= 1) kth = fmin(kth, __shfl_xor_sync(0xffffffffu, kth, o));
  if (lane == 0) {
    bool ok;
    if (sel_n[q] <= kKeep && (long long)sel_n[q] >= h->R) {
      ok = true;                                   // every row of the shard is a candidate and was rescored
    } else {
      const double eps = knn_eps(sqrt(qn2[q]), qexp[q], h);
      ok = kth < double(sel_T[q]) + qn2[q] - eps;
    }
    // two-phase query: the cutoff kernel has already established that every row of this shard that can be in the GLOBAL
    // top-k was rescored (knn_apply_cutoff_kernel); the list may then hold fewer than k rows
    if (cut_mode != nullptr && cut_mode[q]) ok = true;
    if (force_all) ok = false;
    kth_d2[q] = kth;                                // k-th smallest exact d^2 among the rescored candidates (inf: fewer than k)
    if (ok) {
      atomicAdd(&stats[1], 1);
    } else {
      const int slot = atomicAdd(&stats[2], 1);
      flag_list[slot] = q0 + q;                     // all pointers of this launch are chunk-relative; the list is global
    }
  }
}

// Two-phase (sharded) query, between the phases: ub_all [G,Q,k] = the ranks' per-query lists of upper bounds (ascending,
// each the bound of one actual row); bound[q] = the k-th smallest of the G*k values: k distinct rows of the database have an
// exact distance at or below it.  One warp per query; every element finds its rank in the union by binary search in
// the other lists (ties: by list, then position), the one of rank k-1 is the answer.
constexpr int kBoundStage = 512;                     // G * k values per query staged in shared memory (8 ranks x k <= 32: 256)
__global__ void __launch_bounds__(256) knn_bound_reduce_kernel(const float* __restrict__ ub_all, int G, int Q, int k,
                                                               float* __restrict__ bound) {
  __shared__ float s_ub[8][kBoundStage];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= Q) return;
  const bool staged = G * k <= kBoundStage;          // the query's G lists side by side in shared memory
  if (staged) {
    for (int e = lane; e < G * k; e += 32) s_ub[warp][e] = ub_all[(size_t(e / k) * Q + q) * k + (e % k)];
    __syncwarp();
  }
  for (int e = lane; e < G * k; e += 32) {
    const int g = e / k, j = e - g * k;
    const float v = staged ? s_ub[warp][e] : ub_all[(size_t(g) * Q + q) * k + j];
    int rank = j;
    for (int g2 = 0; g2 < G && rank < k; ++g2) {
      if (g2 == g) continue;
      const float* l = staged ? &s_ub[warp][g2 * k] : ub_all + (size_t(g2) * Q + q) * k;
      int lo = 0, hi = k;              // entries of list g2 ordered before (v, g): < v, or == v when g2 < g
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float o = l[mid];
        if (o < v || (o == v && g2 < g)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank == k - 1) bound[q] = v;
  }
}

// Two-phase (sharded) query, second phase.  bound[q] as above (each rank's k best fp16 scores + its eps, k-th smallest
// of the union): at least k rows somewhere have an exact distance <= bound (score units), so a row of THIS shard whose fp16 score exceeds
// cut = bound + eps_here is strictly farther than k rows and cannot be in the global top-k.  If the shard's bound on
// everything that is NOT a candidate (sel_T) also exceeds the cut, the candidates at or below it are all this shard can
// contribute: they alone are rescored (a few per query at 8 shards instead of k..64) and the query is certified here.
// Otherwise (more than kKeep rows of the shard inside the cut) the query keeps every candidate and goes through the
// single-rank certificate / fallback chain, which returns its exact local top-k.
__global__ void __launch_bounds__(256) knn_apply_cutoff_kernel(const float* __restrict__ bound, const float* __restrict__ sel_score,
                                                               const float* __restrict__ sel_T, const int* __restrict__ sel_n,
                                                               const double* __restrict__ qn2, const int* __restrict__ qexp,
                                                               const ShadowHeader* __restrict__ h, int Q,
                                                               uint32_t* __restrict__ sel_idx, int* __restrict__ cut_mode,
                                                               int* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  float cut = INFINITY;
  if (lane == 0) {
    const float b = bound[q];
    if (isfinite(b)) cut = __double2float_ru(double(b) + knn_eps(sqrt(qn2[q]), qexp[q], h));
  }
  cut = __shfl_sync(0xffffffffu, cut, 0);
  const bool all_rows = sel_n[q] <= kKeep && (long long)sel_n[q] >= h->R;    // every row of the shard is a candidate
  const bool mode = isfinite(cut) && (sel_T[q] > cut || all_rows);
  if (mode) {
#pragma unroll
    for (int u = 0; u < kKeep / 32; ++u) {
      const int t = lane + 32 * u;
      if (sel_score[size_t(q) * kKeep + t] > cut) sel_idx[size_t(q) * kKeep + t] = 0xffffffffu;
    }
  }
  if (lane == 0) {
    cut_mode[q] = mode ? 1 : 0;
    if (mode) atomicAdd(&stats[7], 1);
  }
}

// ---------------------------------------------------------------------------------------------
// exact scan: float64 distances of QT staged queries against every row, then radix select
// ---------------------------------------------------------------------------------------------
constexpr int kScanQT = 8;

// qsel == nullptr: queries q0..q0+qt-1; else queries qsel[q0..].  d2 [qt, R].
__global__ void __launch_bounds__(256) knn_scan_kernel(const float* __restrict__ db, long long R, int D,
                                                       const float* __restrict__ q, const int* __restrict__ qsel,
                                                       int q0, int qt, double* __restrict__ d2) {
  extern __shared__ __align__(16) float sq[];   // [qt][D]
  for (int i = threadIdx.x; i < qt * (D >> 2); i += blockDim.x) {
    const int qq = i / (D >> 2), c = i - qq * (D >> 2);
    const int qi = qsel ? qsel[q0 + qq] : q0 + qq;
    reinterpret_cast<float4*>(sq)[i] = __ldg(reinterpret_cast<const float4*>(q + size_t(qi) * D) + c);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp; r < R; r += nw) {
    const float4* row = reinterpret_cast<const float4*>(db + size_t(r) * D);
    double acc[kScanQT];
#pragma unroll
    for (int u = 0; u < kScanQT; ++u) acc[u] = 0.0;
    for (int c = lane; c < (D >> 2); c += 32) {
      const float4 b = ldg_stream(row + c);
      const double bx = b.x, by = b.y, bz = b.z, bw = b.w;
#pragma unroll
      for (int u = 0; u < kScanQT; ++u) {
        if (u < qt) {
          const float4 a = reinterpret_cast<const float4*>(sq)[u * (D >> 2) + c];
          double d;
          d = double(a.x) - bx; acc[u] = fma(d, d, acc[u]);
          d = double(a.y) - by; acc[u] = fma(d, d, acc[u]);
          d = double(a.z) - bz; acc[u] = fma(d, d, acc[u]);
          d = double(a.w) - bw; acc[u] = fma(d, d, acc[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kScanQT; ++u) {
      if (u < qt) {
        const double s = warp_sum(acc[u]);
        if (lane == 0) d2[size_t(u) * R + r] = s;
      }
    }
  }
}

constexpr int kSelThreads = 1024;
constexpr int kTieCap = 4096;

struct SelKey {
  unsigned long long d;   // bit pattern of a non-negative double: orders like the double
  unsigned long long i;
  __device__ bool operator<(const SelKey& o) const { return d < o.d || (d == o.d && i < o.i); }
};

// One CTA per query: MSB-first radix select (8 bits per pass) of the k-th smallest d^2, then gather + sort.
// d2 [nq, R]; out slot given by qsel (or q0 + blockIdx.x).
__global__ void __launch_bounds__(kSelThreads) knn_select_kernel(const double* __restrict__ d2, long long R, int k,
                                                                 const int* __restrict__ qsel, int q0,
                                                                 long long idx_offset, SelKey* __restrict__ scratch,
                                                                 double* __restrict__ out_d,
                                                                 long long* __restrict__ out_i) {
  extern __shared__ unsigned char sel_smem[];
  SelKey* skeys = reinterpret_cast<SelKey*>(sel_smem);     // [kpow2]
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned int s_krem, s_nless, s_neq;
  const int qq = blockIdx.x;
  const int qi = qsel ? qsel[q0 + qq] : q0 + qq;
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(d2 + size_t(qq) * R);
  const int kk = int(min((long long)k, R));
  if (threadIdx.x == 0) { s_prefix = 0ull; s_krem = unsigned(kk); }
  for (int pass = 0; pass < 8; ++pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const int shift = 56 - 8 * pass;
    for (long long r = threadIdx.x; r < R; r += blockDim.x) {
      const unsigned long long key = keys[r];
      const bool match = pass == 0 ? true : ((key >> (shift + 8)) == (prefix >> (shift + 8)));
      if (match) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int krem = s_krem, cum = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (cum + hist[b] >= krem) break;
        cum += hist[b];
      }
      s_krem = krem - cum;
      s_prefix = prefix | ((unsigned long long)b << shift);
    }
    __syncthreads();
  }
  const unsigned long long kth = s_prefix;      // exact bit pattern of the k-th smallest d^2
  const unsigned int need_eq = s_krem;           // how many rows equal to kth belong to the answer
  if (threadIdx.x == 0) { s_nless = 0u; s_neq = 0u; }
  int kpow2 = 1;
  while (kpow2 < kk) kpow2 <<= 1;
  SelKey* ties = scratch + size_t(qq) * kTieCap;
  __syncthreads();
  for (long long r = threadIdx.x; r < R; r += blockDim.x) {
    const unsigned long long key = keys[r];
    if (key < kth) {
      const unsigned int slot = atomicAdd(&s_nless, 1u);
      skeys[slot].d = key;
      skeys[slot].i = (unsigned long long)r;
    } else if (key == kth) {
      const unsigned int slot = atomicAdd(&s_neq, 1u);
      if (slot < kTieCap) { ties[slot].d = key; ties[slot].i = (unsigned long long)r; }
    }
  }
  __syncthreads();
  const unsigned int nless = s_nless, neq = min(s_neq, (unsigned int)kTieCap);
  // among the rows tied at the k-th distance keep the need_eq smallest indices
  if (neq == need_eq) {
    for (unsigned int e = threadIdx.x; e < neq; e += blockDim.x) skeys[nless + e] = ties[e];
  } else {
    for (unsigned int e = threadIdx.x; e < neq; e += blockDim.x) {
      const unsigned long long mine = ties[e].i;
      unsigned int rank = 0;
      for (unsigned int x = 0; x < neq; ++x) rank += ties[x].i < mine ? 1u : 0u;
      if (rank < need_eq) skeys[nless + rank] = ties[e];
    }
  }
  for (int e = int(nless + need_eq) + threadIdx.x; e < kpow2; e += blockDim.x) { skeys[e].d = ~0ull; skeys[e].i = ~0ull; }
  bitonic_sort_smem(skeys, kpow2);
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    if (e < kk) {
      out_d[size_t(qi) * k + e] = sqrt(__longlong_as_double((long long)skeys[e].d));
      out_i[size_t(qi) * k + e] = (long long)skeys[e].i + idx_offset;
    } else {
      out_d[size_t(qi) * k + e] = INFINITY;
      out_i[size_t(qi) * k + e] = -1ll;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// shard merge, geo bookkeeping, recall
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_less(double d0, long long i0, double d1, long long i1) {
  // padding entries (index < 0) sort last
  if ((i0 < 0) != (i1 < 0)) return i1 < 0;
  return d0 < d1 || (d0 == d1 && i0 < i1);
}

// shard g's lists start at d_all + g*stride / i_all + g*stride (elements): lets one packed all-gather buffer
// [G][dist Q*k | idx Q*k] feed the merge without a repack
// One warp per query.  The query's G lists (G*k (distance, index) pairs; 8 ranks x 25: 3.2 KB) are staged in shared
// memory when they fit, so that the G-1 binary searches of every entry run there instead of in global memory; the
// searches cover only the real prefix of a list (a two-phase call returns ~k/G rows per rank, the rest is (inf, -1)
// padding), and padding is ranked only when the lists together hold fewer than k rows.
constexpr int kMergeStage = 256;                     // staged pairs per query (8 warps x 256 x 16 B = 32 KB per CTA)
constexpr int kMergeLists = 64;                      // staged list lengths
__global__ void __launch_bounds__(256) topk_merge_kernel(const double* __restrict__ d_all, const long long* __restrict__ i_all,
                                                         int G, int Q, int k, long long stride, double* __restrict__ d,
                                                         long long* __restrict__ i) {
  __shared__ double s_d[8][kMergeStage];
  __shared__ long long s_i[8][kMergeStage];
  __shared__ int s_real[8][kMergeLists];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 8 + warp;
  if (q >= Q) return;
  const int n = G * k;
  const bool staged = n <= kMergeStage && G <= kMergeLists;
  int total_real = 0;
  if (staged) {
    for (int g = lane; g < G; g += 32) s_real[warp][g] = 0;
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
      const int g = e / k, j = e - g * k;
      s_d[warp][e] = d_all[(size_t)g * stride + (size_t)q * k + j];
      const long long v = i_all[(size_t)g * stride + (size_t)q * k + j];
      s_i[warp][e] = v;
      if (v >= 0) atomicAdd(&s_real[warp][g], 1);
    }
    __syncwarp();
    for (int g = lane; g < G; g += 32) total_real += s_real[warp][g];
    total_real = __reduce_add_sync(0xffffffffu, total_real);
  }
  for (int e = lane; e < n; e += 32) {
    const int g = e / k, j = e - g * k;
    const double md = staged ? s_d[warp][e] : d_all[(size_t)g * stride + (size_t)q * k + j];
    const long long mi = staged ? s_i[warp][e] : i_all[(size_t)g * stride + (size_t)q * k + j];
    if (staged && mi < 0 && total_real >= k) continue;         // padding cannot reach the output
    // rank = number of entries, over all shard lists of this query, that order before (md, mi)
    int rank = j;    // entries before it in its own (sorted) list
    for (int g2 = 0; g2 < G && rank < k; ++g2) {
      if (g2 == g) continue;
      const double* ld = staged ? &s_d[warp][g2 * k] : d_all + (size_t)g2 * stride + (size_t)q * k;
      const long long* li = staged ? &s_i[warp][g2 * k] : i_all + (size_t)g2 * stride + (size_t)q * k;
      // first position whose entry is NOT less than mine; padding (inf, -1) orders after every real entry and ties with
      // other padding, so for a real entry the search stops at the real prefix
      int lo = 0, hi = (staged && mi >= 0) ? s_real[warp][g2] : k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pair_less(ld[mid], li[mid], md, mi)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      d[(size_t)q * k + rank] = mi < 0 ? INFINITY : md;
      i[(size_t)q * k + rank] = mi < 0 ? -1ll : mi;
    }
  }
}

__global__ void __launch_bounds__(256) geo_topn_kernel(const double* __restrict__ qxy, const double* __restrict__ rxy,
                                                       int Q, long long R, const long long* __restrict__ top_i, int k,
                                                       double* __restrict__ top_g, long long* __restrict__ gt_i,
                                                       double* __restrict__ gt_d) {
  __shared__ double s_d[8];
  __shared__ long long s_i[8];
  const int q = blockIdx.x;
  const double x = qxy[2 * q], y = qxy[2 * q + 1];
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const long long r = top_i[(size_t)q * k + j];
    double v = INFINITY;
    if (r >= 0 && r < R) {
      const double dx = x - rxy[2 * r], dy = y - rxy[2 * r + 1];
      v = sqrt(dx * dx + dy * dy);
    }
    top_g[(size_t)q * k + j] = v;
  }
  double best = INFINITY;
  long long bi = -1;
  for (long long r = threadIdx.x; r < R; r += blockDim.x) {
    const double dx = x - rxy[2 * r], dy = y - rxy[2 * r + 1];
    const double v = dx * dx + dy * dy;
    if (v < best) { best = v; bi = r; }      // strided scan keeps the lowest index within a thread
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < best || (ov == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_d[warp] = best; s_i[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w)
      if (s_d[w] < best || (s_d[w] == best && s_i[w] >= 0 && (bi < 0 || s_i[w] < bi))) { best = s_d[w]; bi = s_i[w]; }
    gt_i[q] = bi;                              // np.argmin: first minimum
    gt_d[q] = sqrt(best);
  }
}

__global__ void recall_zero_kernel(double* curves, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) curves[i] = 0.0;
}
__global__ void __launch_bounds__(256) recall_count_kernel(const double* __restrict__ top_g, int Q, int k,
                                                           const double* __restrict__ thr, int nx,
                                                           double* __restrict__ curves) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  double m = INFINITY;
  for (int n = 0; n < k; ++n) {
    m = fmin(m, top_g[(size_t)q * k + n]);       // top_n[q,n] = min(d[q,0..n])   (train.py:368-371)
    for (int x = 0; x < nx; ++x)
      if (m < thr[x]) atomicAdd(&curves[n * nx + x], 1.0);    // counts are small integers: exact in float64
  }
}
__global__ void recall_scale_kernel(double* curves, int n, int Q) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) curves[i] = curves[i] / double(Q) * 100.0;       // float(sum(...)) / float(len) * 100  (roc.py:216)
}

// ---------------------------------------------------------------------------------------------
// stage 2 (tensor pass, "collect" mode): queries the first pass could not certify
// ---------------------------------------------------------------------------------------------
// The first pass refuses a query when more than k' = 64 rows lie within the fp16 rounding bound of its k-th neighbour
// (near-duplicate descriptors: consecutive frames of a standing vehicle).  Its k rescored candidates still give an
// UPPER bound e_k on the exact k-th squared distance, and every true top-k row r satisfies
//     score16(r) <= exact(r) + eps <= (e_k - |q|^2) + eps.
// Stage 2 reruns the tensor pass for the refused queries only, with that fixed threshold, and keeps EVERY row below it
// (up to kCollectCap); all of them are rescored exactly and the k best are emitted.  Exact unless the list overflows
// (more than kCollectCap rows inside the bound), in which case the query goes to the float64 scan.
__global__ void __launch_bounds__(256) knn_stage2_gather_kernel(const int* __restrict__ flag_list, int f0, int n2, int Dp,
                                                                const __half* __restrict__ qh, const float* __restrict__ qmul,
                                                                const double* __restrict__ qn2, const int* __restrict__ qexp,
                                                                const double* __restrict__ kth_d2,
                                                                const ShadowHeader* __restrict__ h, __half* __restrict__ qh2,
                                                                float* __restrict__ qmul2, float* __restrict__ thr2,
                                                                int* __restrict__ cnt2) {
  const int f = blockIdx.x;
  if (f >= n2) return;
  const int q = flag_list[f0 + f];
  const uint4* src = reinterpret_cast<const uint4*>(qh + size_t(q) * Dp);
  uint4* dst = reinterpret_cast<uint4*>(qh2 + size_t(f) * Dp);
  for (int c = threadIdx.x; c < (Dp >> 3); c += blockDim.x) dst[c] = src[c];
  if (threadIdx.x == 0) {
    qmul2[f] = qmul[q];
    cnt2[f] = 0;
    const double bound = kth_d2[q] - qn2[q] + knn_eps(sqrt(qn2[q]), qexp[q], h);
    // the kernel tests score < thr: round the bound up and step once more so that equality passes
    float t = isfinite(bound) ? __double2float_ru(bound) : INFINITY;
    if (isfinite(t)) t = nextafterf(t, INFINITY);
    thr2[f] = t;
  }
}

// exact squared distances of the collected rows: one warp per (refused query, list slot); same arithmetic as
// knn_rescore_kernel / knn_scan_kernel (bit-identical d^2)
__global__ void __launch_bounds__(256) knn_stage2_rescore_kernel(const float* __restrict__ db, const float* __restrict__ q,
                                                                 int D, const int* __restrict__ flag_list, int f0,
                                                                 const uint32_t* __restrict__ coll_idx,
                                                                 const int* __restrict__ cnt2, long long pairs,
                                                                 double* __restrict__ d2) {
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= pairs) return;
  const int f = int(p / kCollectCap), e = int(p - (long long)f * kCollectCap);
  const int n = cnt2[f];
  if (n > kCollectCap || e >= n) return;
  const int qi = flag_list[f0 + f];
  const uint32_t r = coll_idx[p];
  const float4* qr = reinterpret_cast<const float4*>(q + size_t(qi) * D);
  const float4* rr = reinterpret_cast<const float4*>(db + size_t(r) * D);
  double acc = 0.0;
  for (int c = lane; c < (D >> 2); c += 32) {
    const float4 a = __ldg(qr + c);
    const float4 b = ldg_stream(rr + c);
    double d;
    d = double(a.x) - double(b.x); acc = fma(d, d, acc);
    d = double(a.y) - double(b.y); acc = fma(d, d, acc);
    d = double(a.z) - double(b.z); acc = fma(d, d, acc);
    d = double(a.w) - double(b.w); acc = fma(d, d, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) d2[p] = acc;
}

// one CTA per refused query: order the collected rows by (d^2, idx), emit the top-k; overflowed lists go to the scan
__global__ void __launch_bounds__(256) knn_stage2_select_kernel(const int* __restrict__ flag_list, int f0,
                                                                const uint32_t* __restrict__ coll_idx,
                                                                const int* __restrict__ cnt2, const double* __restrict__ d2,
                                                                int k, long long idx_offset, double* __restrict__ out_d,
                                                                long long* __restrict__ out_i, int* __restrict__ flag2_list,
                                                                int* __restrict__ stats) {
  extern __shared__ unsigned char sel2_smem[];
  SelKey* keys = reinterpret_cast<SelKey*>(sel2_smem);      // [kCollectCap]
  const int f = blockIdx.x;
  const int q = flag_list[f0 + f];
  const int n = cnt2[f];
  if (n > kCollectCap || n < k) {                            // overflow (or an unusable bound): exact scan
    if (threadIdx.x == 0) flag2_list[atomicAdd(&stats[5], 1)] = q;
    return;
  }
  int n_pow2 = 32;
  while (n_pow2 < n) n_pow2 <<= 1;
  for (int e = threadIdx.x; e < n_pow2; e += blockDim.x) {
    SelKey key;
    if (e < n) {
      key.d = (unsigned long long)__double_as_longlong(d2[size_t(f) * kCollectCap + e]);
      key.i = coll_idx[size_t(f) * kCollectCap + e];
    } else {
      key.d = ~0ull; key.i = ~0ull;
    }
    keys[e] = key;
  }
  bitonic_sort_smem(keys, n_pow2);
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    out_d[size_t(q) * k + e] = sqrt(__longlong_as_double((long long)keys[e].d));
    out_i[size_t(q) * k + e] = (long long)keys[e].i + idx_offset;
  }
  if (threadIdx.x == 0) atomicAdd(&stats[4], 1);
}

// stats out: {n_queries, n_certified, n_refused by the first pass, path, n_stage2, n_scan, chunks, 0}
struct StatsOut { int v[8]; };
__global__ void knn_stats_out_kernel(int* __restrict__ out, StatsOut s) {
  if (threadIdx.x < 8) out[threadIdx.x] = s.v[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------------
// Measurement hook state (scl_knn_timing): process-wide, guarded by a mutex.
static std::atomic<int> g_knn_timing{0};
static std::mutex g_knn_timing_mu;
static double g_knn_tc_ms_sum = 0.0;
static int g_knn_tc_calls = 0;

// test hook (scl_knn_set_debug_scores): per host thread, size-checked
static thread_local float* t_dbg_scores = nullptr;
static thread_local size_t t_dbg_capacity = 0;

// Per host thread and device: the helper stream on which the candidate merge / exact rescore / certificate of query chunk
// i run while the tensor pass of chunk i+1 occupies the launching stream, and the events that order the two.
struct HelperCtx {
  cudaStream_t h = nullptr;
  cudaEvent_t armed = nullptr;          // scl_knn_query_launch: group counters zeroed (consumer streams wait on it)
  std::vector<cudaEvent_t> ev;          // ordering events (timing disabled)
  std::vector<cudaEvent_t> tev;         // timing events (scl_knn_timing)
  int get(std::vector<cudaEvent_t>& pool, size_t i, unsigned flags, cudaEvent_t* out) {
    while (pool.size() <= i) {
      cudaEvent_t e;
      SCL_CUDA_TRY(cudaEventCreateWithFlags(&e, flags));
      pool.push_back(e);
    }
    *out = pool[i];
    return SCL_OK;
  }
};
static thread_local HelperCtx t_helper[kMaxDevices];

struct QueryWs {
  // per query (all Q)
  __half* qh;
  float* qmul;
  double* qn2;
  int* qexp;
  unsigned int* q_thr;
  uint32_t* sel_idx;
  float* sel_T;
  int* sel_n;
  double* d2;
  double* kth_d2;
  float* sel_score;
  int* cut_mode;
  int* flag_list;
  int* flag2_list;
  int* stats;
  // per pipeline slot (two chunks in flight)
  float* cand_s[2];
  uint32_t* cand_i[2];
  int* cand_cnt[2];
  unsigned int* sync_ctr[2];
  unsigned int* sync_ctr2;      // pacing counters of the stage-2 launches (may run while a first pass is still in flight)
  unsigned int* group_done;     // [kMaxGroups] completion counters of the first pass (TcArgs::group_done)
  // stage 2
  __half* qh2;
  float* qmul2;
  float* thr2;
  int* cnt2;
  uint32_t* coll_idx;
  double* coll_d2;
  // exact scan
  double* scan_d2;
  SelKey* sel_scratch;
  int scan_qb;
  int chunk_q, nchunks, nr_max, q2max;
};

static int scan_batch(int64_t R) {
  // queries per exact-scan launch: keep the [qb, R] float64 buffer <= 256 MB
  long long qb = (256ll << 20) / (8ll * (R > 0 ? R : 1));
  if (qb < 1) qb = 1;
  if (qb > 64) qb = 64;
  return int(qb);
}
static int scan_qt(int D) {
  int qt = int((96 * 1024) / (size_t(D) * 4));
  if (qt > kScanQT) qt = kScanQT;
  if (qt < 1) qt = 1;
  return qt;
}

static bool use_tensor_pass(int64_t R, int D, int Q, int k, int force_path) {
  if (force_path == 1) return false;
  if (k > kKeep / 2 || (D & 3) || R >= (1ll << 31) || R < 1024) return false;   // k' = 64 must leave slack above k
  if (force_path >= 2) return true;
  return double(Q) * double(R) >= double(1 << 22) && R >= 4096;
}

// Queries per pipelined chunk: the tensor kernel's own L2 group (knn_tc_tiling: ~40 MB of fp16 query blocks, 5120 queries
// at D = 4096), so that splitting a call at chunk boundaries does not change how the database ranges are streamed.
static int chunk_queries(int Q, int64_t R, int Dp) {
  int forced = knob(KNOB_KNN_CHUNK_Q);
  if (forced == 0) return Q;
  int mb, nt, NR, tpr, gm;
  knn_tc_tiling(Q, R, Dp, &mb, &nt, &NR, &tpr, &gm);
  const int unit = knob_or(KNOB_KNN_TC_VARIANT, 2) >= 2 ? 256 : 128;
  long long cq = forced > 0 ? (long long)(forced + unit - 1) / unit * unit : (long long)gm * unit;
  if (cq < unit) cq = unit;
  if (cq >= Q) return Q;
  // equal-sized chunks (multiples of the tile unit)
  const int n = int((Q + cq - 1) / cq);
  cq = ((Q + n - 1) / n + unit - 1) / unit * unit;
  return int(cq);
}

// one_chunk: the layout of the two-phase (sharded) query, whose first phase has nothing to pipeline the chunks against and
// runs ONE tensor launch over all queries; scl_knn_query_workspace_bytes returns the larger of the two layouts.
static size_t query_ws_layout(int64_t R, int D, int Q, int k, QueryWs* w, void* base, size_t bytes, bool one_chunk = false) {
  Carver c(base, bytes);
  const int Dp = pad64(D);
  QueryWs tmp;
  QueryWs* o = w ? w : &tmp;
  o->chunk_q = one_chunk ? Q : chunk_queries(Q, R, Dp);
  o->nchunks = (Q + o->chunk_q - 1) / o->chunk_q;
  int mb, nt, NR, tpr, gm;
  knn_tc_tiling(o->chunk_q, R, Dp, &mb, &nt, &NR, &tpr, &gm);
  o->nr_max = NR;
  const int q_last = Q - (o->nchunks - 1) * o->chunk_q;
  knn_tc_tiling(q_last, R, Dp, &mb, &nt, &NR, &tpr, &gm);
  if (NR > o->nr_max) o->nr_max = NR;
  o->q2max = Q < 2048 ? Q : 2048;
  knn_tc_tiling(o->q2max, R, Dp, &mb, &nt, &NR, &tpr, &gm);      // stage 2 runs in batches of <= q2max queries
  o->stats = c.take<int>(8);
  o->qh = c.take<__half>(size_t(Q) * Dp);
  o->qmul = c.take<float>(Q);
  o->qn2 = c.take<double>(Q);
  o->qexp = c.take<int>(Q);
  o->q_thr = c.take<unsigned int>(Q);
  o->sel_idx = c.take<uint32_t>(size_t(Q) * kKeep);
  o->sel_T = c.take<float>(Q);
  o->sel_n = c.take<int>(Q);
  o->d2 = c.take<double>(size_t(Q) * kKeep);
  o->kth_d2 = c.take<double>(Q);
  o->sel_score = c.take<float>(size_t(Q) * kKeep);
  o->cut_mode = c.take<int>(Q);
  o->flag_list = c.take<int>(Q);
  o->flag2_list = c.take<int>(Q);
  const int slots = o->nchunks > 1 ? 2 : 1;
  for (int sl = 0; sl < 2; ++sl) {
    if (sl < slots) {
      o->cand_s[sl] = c.take<float>(size_t(o->chunk_q) * o->nr_max * kCandCap);
      o->cand_i[sl] = c.take<uint32_t>(size_t(o->chunk_q) * o->nr_max * kCandCap);
      o->cand_cnt[sl] = c.take<int>(size_t(o->chunk_q) * o->nr_max);
      o->sync_ctr[sl] = c.take<unsigned int>(kSyncMax);
    } else {
      o->cand_s[sl] = o->cand_s[0]; o->cand_i[sl] = o->cand_i[0]; o->cand_cnt[sl] = o->cand_cnt[0];
      o->sync_ctr[sl] = o->sync_ctr[0];
    }
  }
  o->sync_ctr2 = c.take<unsigned int>(kSyncMax);
  o->group_done = c.take<unsigned int>(kMaxGroups);
  o->qh2 = c.take<__half>(size_t(o->q2max) * Dp);
  o->qmul2 = c.take<float>(o->q2max);
  o->thr2 = c.take<float>(o->q2max);
  o->cnt2 = c.take<int>(o->q2max);
  o->coll_idx = c.take<uint32_t>(size_t(o->q2max) * kCollectCap);
  o->coll_d2 = c.take<double>(size_t(o->q2max) * kCollectCap);
  o->scan_qb = scan_batch(R);
  o->scan_d2 = c.take<double>(size_t(o->scan_qb) * R);
  o->sel_scratch = c.take<SelKey>(size_t(o->scan_qb) * kTieCap);
  (void)k;
  return c.off;
}

// exact scan + select for `nq` queries (all of them, or those listed in qsel)
static int run_exact(const float* db, int64_t R, int D, const float* queries, const int* qsel, int nq, int k,
                     int64_t idx_offset, double* dist, int64_t* idx, const QueryWs& w, cudaStream_t stream) {
  const int qt = scan_qt(D);
  const size_t scan_smem = size_t(qt) * D * sizeof(float);
  static SmemAttrCache scan_cfg;                        // per device
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(knn_scan_kernel), scan_smem, &scan_cfg);
  if (rc) return rc;
  int kpow2 = 1;
  const int kk = int(k < R ? k : R);
  while (kpow2 < kk) kpow2 <<= 1;
  const size_t sel_smem = size_t(kpow2) * sizeof(SelKey);
  long long rows_per_cta = 8;
  int scan_grid = int(std::min<long long>((R + rows_per_cta - 1) / rows_per_cta, (long long)num_sms() * 8));
  for (int b0 = 0; b0 < nq; b0 += w.scan_qb) {
    const int nb = std::min(w.scan_qb, nq - b0);
    for (int s0 = 0; s0 < nb; s0 += qt) {
      const int n = std::min(qt, nb - s0);
      knn_scan_kernel<<<scan_grid, 256, scan_smem, stream>>>(db, R, D, queries, qsel, b0 + s0, n,
                                                             w.scan_d2 + size_t(s0) * R);
      SCL_LAUNCH_CHECK();
    }
    knn_select_kernel<<<nb, kSelThreads, sel_smem, stream>>>(w.scan_d2, R, k, qsel, b0, idx_offset, w.sel_scratch, dist,
                                                             reinterpret_cast<long long*>(idx));
    SCL_LAUNCH_CHECK();
  }
  return SCL_OK;
}

// tiling + launch of the tensor pass for `nq` queries whose per-query arrays start at qh / qmul / q_thr
static int launch_tensor(const QueryWs& w, int slot, const __half* qh, const float* qmul, unsigned int* q_thr, int nq,
                         int64_t R, int Dp, const float* rn, const __half* dbh, float* dbg, bool collect,
                         cudaStream_t stream, bool signal_groups = false) {
  TcArgs a = {};
  a.rn = rn; a.qmul = qmul; a.Q = nq; a.R = int(R); a.Dp = Dp;
  knn_tc_tiling(nq, R, Dp, &a.num_m_blocks, &a.num_n_tiles, &a.NR, &a.tiles_per_range, &a.group_m);
  a.cand_s = w.cand_s[slot]; a.cand_i = w.cand_i[slot]; a.cand_cnt = w.cand_cnt[slot]; a.q_thr = q_thr;
  a.sync_ctr = collect ? w.sync_ctr2 : w.sync_ctr[slot];
  SCL_CUDA_TRY(cudaMemsetAsync(a.sync_ctr, 0, size_t(kSyncMax) * sizeof(unsigned int), stream));
  if (signal_groups) a.group_done = w.group_done;      // zeroed by the caller
  a.dbg_scores = dbg;
  if (collect) {
    a.collect = 1; a.fixed_thr = w.thr2; a.coll_idx = w.coll_idx; a.coll_cnt = w.cnt2; a.coll_cap = kCollectCap;
  }
  return knn_tc_launch(a, qh, dbh, stream);
}

}  // namespace scl

using namespace scl;

// Queries the first pass refused (hs = host copy of the counters, read after the stream was synchronised): second tensor
// stage, then the exact scan for what is left; finally the counters go out.  Shared by scl_knn_query and scl_knn_query_end.
static int resolve_refused(int (&hs)[8], const float* db, int64_t R, int D, const float* queries, int Q, int k,
                           int64_t idx_offset, int force_path, double* dist, int64_t* idx, int32_t* stats, const QueryWs& w,
                           const ShadowHeader* h, const float* rn, const __half* dbh, int Dp, int nchunks,
                           cudaStream_t stream) {
  int rc;
  const int nflag = hs[2];
  int n_stage2 = 0, n_scan = 0;
  if (nflag > 0) {
    // force_path 3: every query through the exact scan; 4: every query through stage 2 (test hooks)
    const bool stage2 = force_path != 3 && knob_or(KNOB_KNN_STAGE2, 1) != 0;
    if (!stage2) {
      rc = run_exact(db, R, D, queries, w.flag_list, nflag, k, idx_offset, dist, idx, w, stream);
      if (rc) return rc;
      n_scan = nflag;
    } else {
      static SmemAttrCache sel2_cfg;
      constexpr size_t sel2_smem = size_t(kCollectCap) * sizeof(SelKey);
      if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(knn_stage2_select_kernel), sel2_smem, &sel2_cfg))) return rc;
      for (int f0 = 0; f0 < nflag; f0 += w.q2max) {
        const int n2 = std::min(w.q2max, nflag - f0);
        knn_stage2_gather_kernel<<<n2, 256, 0, stream>>>(w.flag_list, f0, n2, Dp, w.qh, w.qmul, w.qn2, w.qexp, w.kth_d2, h,
                                                         w.qh2, w.qmul2, w.thr2, w.cnt2);
        SCL_LAUNCH_CHECK();
        rc = launch_tensor(w, 0, w.qh2, w.qmul2, w.q_thr, n2, R, Dp, rn, dbh, nullptr, true, stream);
        if (rc) return rc;
        const long long pairs = (long long)n2 * kCollectCap;
        knn_stage2_rescore_kernel<<<unsigned((pairs + 7) / 8), 256, 0, stream>>>(db, queries, D, w.flag_list, f0, w.coll_idx,
                                                                                w.cnt2, pairs, w.coll_d2);
        SCL_LAUNCH_CHECK();
        knn_stage2_select_kernel<<<n2, 256, sel2_smem, stream>>>(w.flag_list, f0, w.coll_idx, w.cnt2, w.coll_d2, k,
                                                                 idx_offset, dist, reinterpret_cast<long long*>(idx),
                                                                 w.flag2_list, w.stats);
        SCL_LAUNCH_CHECK();
      }
      // how many lists overflowed decides whether the exact scan runs at all: second (and last) host read
      SCL_CUDA_TRY(cudaMemcpyAsync(hs, w.stats, sizeof(hs), cudaMemcpyDeviceToHost, stream));
      SCL_CUDA_TRY(cudaStreamSynchronize(stream));
      n_stage2 = hs[4];
      n_scan = hs[5];
      if (n_scan > 0) {
        rc = run_exact(db, R, D, queries, w.flag2_list, n_scan, k, idx_offset, dist, idx, w, stream);
        if (rc) return rc;
      }
    }
  }
  if (stats) {
    const StatsOut so = {{Q, hs[1], nflag, 2, n_stage2, n_scan, nchunks, hs[7]}};
    knn_stats_out_kernel<<<1, 8, 0, stream>>>(stats, so);
    SCL_LAUNCH_CHECK();
  }
  return SCL_OK;
}


extern "C" int scl_knn_shadow_bytes(int64_t R, int D, size_t* bytes) {
  if (!bytes || R < 1 || D < 4 || (D & 3)) return SCL_ERR_BAD_ARG;
  *bytes = shadow_data_off(R) + size_t(R) * pad64(D) * sizeof(__half);
  return SCL_OK;
}

extern "C" int scl_knn_build(const float* db, int64_t R, int D, void* shadow, size_t shadow_bytes, scl_stream_t stream_) {
  if (!db || !shadow || R < 1) return SCL_ERR_BAD_ARG;
  if (D < 4 || (D & 3)) return SCL_ERR_BAD_SHAPE;
  if (!aligned16(db) || (reinterpret_cast<uintptr_t>(shadow) & 255u)) return SCL_ERR_ALIGN;
  size_t need = 0;
  scl_knn_shadow_bytes(R, D, &need);
  if (shadow_bytes < need) return SCL_ERR_WORKSPACE;
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  char* sb = static_cast<char*>(shadow);
  ShadowHeader* h = reinterpret_cast<ShadowHeader*>(sb);
  float* rn = reinterpret_cast<float*>(sb + shadow_norm_off());
  __half* data = reinterpret_cast<__half*>(sb + shadow_data_off(R));
  const int Dp = pad64(D);
  knn_header_init_kernel<<<1, 1, 0, stream>>>(h, R, D, Dp);
  SCL_LAUNCH_CHECK();
  const int grid = num_sms() * 8;
  knn_build_stats_kernel<<<grid, 256, 0, stream>>>(db, R, D, rn, h);
  SCL_LAUNCH_CHECK();
  knn_build_scale_kernel<<<1, 1, 0, stream>>>(h);
  SCL_LAUNCH_CHECK();
  knn_build_convert_kernel<<<grid, 256, 0, stream>>>(db, R, D, Dp, h, data);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

extern "C" int scl_knn_query_workspace_bytes(int64_t R, int D, int Q, int k, size_t* bytes) {
  if (!bytes || R < 1 || Q < 1 || k < 1 || D < 4 || (D & 3)) return SCL_ERR_BAD_ARG;
  *bytes = std::max(query_ws_layout(R, D, Q, k, nullptr, nullptr, 0), query_ws_layout(R, D, Q, k, nullptr, nullptr, 0, true));
  return SCL_OK;
}

extern "C" int scl_knn_set_debug_scores(float* scores, size_t capacity_floats) {
  t_dbg_scores = scores;
  t_dbg_capacity = scores ? capacity_floats : 0;
  return SCL_OK;
}

extern "C" int scl_knn_query(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                             int64_t idx_offset, int force_path, double* dist, int64_t* idx, int32_t* stats,
                             void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!db || !queries || !dist || !idx || !workspace) return SCL_ERR_BAD_ARG;
  if (R < 1 || Q < 1 || k < 1 || k > 1024 || D < 4 || (D & 3)) return SCL_ERR_BAD_SHAPE;
  if (!aligned16(db) || !aligned16(queries) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return SCL_ERR_ALIGN;
  int rc = check_device();
  if (rc) return rc;
  QueryWs w;
  const size_t need = query_ws_layout(R, D, Q, k, &w, workspace, workspace_bytes);
  if (workspace_bytes < need) return SCL_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool tensor = use_tensor_pass(R, D, Q, k, force_path);
  if (tensor && !shadow) return SCL_ERR_BAD_ARG;
  SCL_CUDA_TRY(cudaMemsetAsync(w.stats, 0, 8 * sizeof(int), stream));

  if (!tensor) {
    rc = run_exact(db, R, D, queries, nullptr, Q, k, idx_offset, dist, idx, w, stream);
    if (rc) return rc;
    if (stats) {
      const StatsOut so = {{Q, 0, Q, 1, 0, Q, 0, 0}};
      knn_stats_out_kernel<<<1, 8, 0, stream>>>(stats, so);
      SCL_LAUNCH_CHECK();
    }
    return SCL_OK;
  }

  const char* sb = static_cast<const char*>(shadow);
  const ShadowHeader* h = reinterpret_cast<const ShadowHeader*>(sb);
  const float* rn = reinterpret_cast<const float*>(sb + shadow_norm_off());
  const __half* dbh = reinterpret_cast<const __half*>(sb + shadow_data_off(R));
  const int Dp = pad64(D);

  knn_query_prep_kernel<<<Q, 256, 0, stream>>>(queries, Q, D, Dp, h, w.qh, w.qmul, w.qn2, w.qexp);
  SCL_LAUNCH_CHECK();
  SCL_CUDA_TRY(cudaMemsetAsync(w.q_thr, 0xff, size_t(Q) * sizeof(unsigned int), stream));

  // Pipeline over query chunks: the tensor pass of chunk c runs on the launching stream while the candidate merge, the
  // exact rescore and the certificate of chunk c-1 run on the helper stream (the tensor kernel leaves registers and
  // ~20 KB of shared memory per SM free and is tensor-bound; the rescore is an HBM gather).  Two candidate-list slots.
  HelperCtx& hc = t_helper[device_slot()];
  const int nchunks = w.nchunks;
  if (nchunks > 1 && !hc.h) SCL_CUDA_TRY(cudaStreamCreateWithFlags(&hc.h, cudaStreamNonBlocking));
  const bool timing = g_knn_timing.load(std::memory_order_relaxed) != 0;
  float* dbg = (t_dbg_scores && t_dbg_capacity >= size_t(Q) * size_t(R)) ? t_dbg_scores : nullptr;
  for (int c = 0; c < nchunks; ++c) {
    const int q0 = c * w.chunk_q, nq = std::min(w.chunk_q, Q - q0), slot = c & 1;
    cudaEvent_t ev_tc = nullptr, ev_post = nullptr, t0 = nullptr, t1 = nullptr;
    if (nchunks > 1) {
      if ((rc = hc.get(hc.ev, size_t(2 * c), cudaEventDisableTiming, &ev_tc))) return rc;
      if ((rc = hc.get(hc.ev, size_t(2 * c + 1), cudaEventDisableTiming, &ev_post))) return rc;
      if (c >= 2) SCL_CUDA_TRY(cudaStreamWaitEvent(stream, hc.ev[2 * (c - 2) + 1], 0));   // the slot's previous reader
    }
    if (timing) {
      if ((rc = hc.get(hc.tev, size_t(2 * c), cudaEventDefault, &t0))) return rc;
      if ((rc = hc.get(hc.tev, size_t(2 * c + 1), cudaEventDefault, &t1))) return rc;
      SCL_CUDA_TRY(cudaEventRecord(t0, stream));
    }
    rc = launch_tensor(w, slot, w.qh + size_t(q0) * Dp, w.qmul + q0, w.q_thr + q0, nq, R, Dp, rn, dbh,
                       dbg ? dbg + size_t(q0) * size_t(R) : nullptr, false, stream);
    if (rc) return rc;
    if (timing) SCL_CUDA_TRY(cudaEventRecord(t1, stream));
    cudaStream_t ps = stream;
    if (nchunks > 1) {
      SCL_CUDA_TRY(cudaEventRecord(ev_tc, stream));
      SCL_CUDA_TRY(cudaStreamWaitEvent(hc.h, ev_tc, 0));
      ps = hc.h;
    }
    int mb, nt, NR, tpr, gm;
    knn_tc_tiling(nq, R, Dp, &mb, &nt, &NR, &tpr, &gm);
    knn_cand_merge_kernel<<<nq, 256, 0, ps>>>(w.cand_s[slot], w.cand_i[slot], w.cand_cnt[slot], w.q_thr + q0, NR, k,
                                              w.qn2 + q0, w.qexp + q0, h, w.sel_idx + size_t(q0) * kKeep, w.sel_T + q0,
                                              w.sel_n + q0, nullptr, nullptr);
    SCL_LAUNCH_CHECK();
    knn_rescore_kernel<<<nq, 256, 0, ps>>>(db, queries + size_t(q0) * D, D, w.sel_idx + size_t(q0) * kKeep,
                                           w.d2 + size_t(q0) * kKeep);
    SCL_LAUNCH_CHECK();
    knn_finalize_kernel<<<(nq + 7) / 8, 256, 0, ps>>>(w.sel_idx + size_t(q0) * kKeep, w.d2 + size_t(q0) * kKeep, w.sel_T + q0,
                                                      w.sel_n + q0, w.qn2 + q0, w.qexp + q0, h, nq, k, idx_offset,
                                                      force_path >= 3 ? 1 : 0, q0, dist + size_t(q0) * k,
                                                      reinterpret_cast<long long*>(idx) + size_t(q0) * k, w.kth_d2 + q0,
                                                      w.flag_list, w.stats, nullptr);
    SCL_LAUNCH_CHECK();
    if (nchunks > 1) SCL_CUDA_TRY(cudaEventRecord(ev_post, hc.h));
  }
  if (nchunks > 1) SCL_CUDA_TRY(cudaStreamWaitEvent(stream, hc.ev[2 * (nchunks - 1) + 1], 0));   // join (the helper is in order)

  // the number of refused queries decides how much work follows: one small device->host read (the only host
  // synchronisation of a fully certified call)
  int hs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  SCL_CUDA_TRY(cudaMemcpyAsync(hs, w.stats, sizeof(hs), cudaMemcpyDeviceToHost, stream));
  SCL_CUDA_TRY(cudaStreamSynchronize(stream));
  if (timing) {
    double sum = 0.0;
    for (int c = 0; c < nchunks; ++c) {
      float ms = 0.0f;
      SCL_CUDA_TRY(cudaEventElapsedTime(&ms, hc.tev[2 * c], hc.tev[2 * c + 1]));
      sum += double(ms);
    }
    std::lock_guard<std::mutex> lk(g_knn_timing_mu);
    g_knn_tc_ms_sum += sum;
    g_knn_tc_calls += 1;
  }
  return resolve_refused(hs, db, R, D, queries, Q, k, idx_offset, force_path, dist, idx, stats, w, h, rn, dbh, Dp, nchunks, stream);
}

// ---------------------------------------------------------------------------------------------
// Sharded retrieval in two phases: between them the ranks exchange the score bounds of their k best
// candidates, so that each rank rescoring exactly only the candidates that can still enter the GLOBAL top-k -- ~k/G of
// them, not the k..64 its own top-k would need.  The first phase is ONE persistent tensor launch over all queries; it
// signals the completion of every query group (TcArgs::group_done), and the per-group entry points below wait for that
// signal ON THEIR STREAM (cuStreamWaitValue32): launched on a second stream, the merge / exchange / rescore / shard merge
// of group g run while the tensor kernel streams the database for group g+1.
static int two_phase_ok(int64_t R, int D, int Q, int k) {
  return use_tensor_pass(R, D, Q, k, 2) ? SCL_OK : SCL_ERR_UNSUPPORTED;
}

struct Groups {
  int n, group_q;          // number of query groups, queries per group (the last may hold fewer)
  int unit_q, group_m;     // queries per work unit, units per group
  int arrivals_per_item;   // epilogue warps that signal one work item
};
// depends on Q and D only (knn_tc_tiling's group_m does not look at R): every rank of a sharded call sees the same groups
static Groups group_layout(int Q, int Dp) {
  int mb, nt, NR, tpr, gm;
  knn_tc_tiling(Q, 1 << 20, Dp, &mb, &nt, &NR, &tpr, &gm);
  const bool pair = knob_or(KNOB_KNN_TC_VARIANT, 2) >= 2;
  const int mu = pair ? (mb + 1) / 2 : mb;
  Groups g;
  g.unit_q = pair ? 256 : 128;
  g.group_m = gm;
  g.n = (mu + gm - 1) / gm;
  g.group_q = gm * g.unit_q;
  g.arrivals_per_item = pair ? 8 : 4;
  return g;
}

typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static int stream_wait_geq(cudaStream_t stream, const unsigned int* addr, unsigned int value) {
  static StreamWaitValue32Fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SCL_CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) {
      set_last_error("cuStreamWaitValue32 entry point", cudaErrorUnknown);
      return SCL_ERR_CUDA;
    }
    fn = reinterpret_cast<StreamWaitValue32Fn>(p);
  }
  const CUresult r = fn(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuStreamWaitValue32", cudaErrorUnknown);
    return SCL_ERR_CUDA;
  }
  return SCL_OK;
}

struct ShardView {
  const ShadowHeader* h;
  const float* rn;
  const __half* dbh;
  int Dp;
};
static ShardView shard_view(const void* shadow, int64_t R, int D) {
  const char* sb = static_cast<const char*>(shadow);
  return {reinterpret_cast<const ShadowHeader*>(sb), reinterpret_cast<const float*>(sb + shadow_norm_off()),
          reinterpret_cast<const __half*>(sb + shadow_data_off(R)), pad64(D)};
}

// query range of a group (group < 0: all queries)
static void group_range(const Groups& g, int Q, int group, int* q0, int* nq) {
  if (group < 0) { *q0 = 0; *nq = Q; return; }
  *q0 = group * g.group_q;
  *nq = std::min(g.group_q, Q - *q0);
}

static int two_phase_args(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                          void* workspace, size_t workspace_bytes, QueryWs* w) {
  if (!db || !shadow || !queries || !workspace) return SCL_ERR_BAD_ARG;
  if (R < 1 || Q < 1 || k < 1 || D < 4 || (D & 3)) return SCL_ERR_BAD_SHAPE;
  if (!aligned16(db) || !aligned16(queries) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return SCL_ERR_ALIGN;
  int rc = two_phase_ok(R, D, Q, k);
  if (rc) return rc;
  if (group_layout(Q, pad64(D)).n > kMaxGroups) return SCL_ERR_UNSUPPORTED;     // > 64 x ~5120 queries at D = 4096: split the call
  if ((rc = check_device())) return rc;
  const size_t need = query_ws_layout(R, D, Q, k, w, workspace, workspace_bytes, true);
  return workspace_bytes < need ? SCL_ERR_WORKSPACE : SCL_OK;
}

extern "C" int scl_knn_query_groups(int D, int Q, int* n_groups, int* group_queries) {
  if (!n_groups || !group_queries || Q < 1 || D < 4 || (D & 3)) return SCL_ERR_BAD_ARG;
  const Groups g = group_layout(Q, pad64(D));
  *n_groups = g.n;
  *group_queries = g.n == 1 ? Q : g.group_q;
  return SCL_OK;
}

extern "C" int scl_knn_query_launch(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                                    void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  QueryWs w;
  int rc = two_phase_args(db, shadow, R, D, queries, Q, k, workspace, workspace_bytes, &w);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const ShardView sv = shard_view(shadow, R, D);
  knn_query_prep_kernel<<<Q, 256, 0, stream>>>(queries, Q, D, sv.Dp, sv.h, w.qh, w.qmul, w.qn2, w.qexp);
  SCL_LAUNCH_CHECK();
  SCL_CUDA_TRY(cudaMemsetAsync(w.q_thr, 0xff, size_t(Q) * sizeof(unsigned int), stream));
  HelperCtx& hc = t_helper[device_slot()];
  const bool timing = g_knn_timing.load(std::memory_order_relaxed) != 0;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (timing) {
    if ((rc = hc.get(hc.tev, 0, cudaEventDefault, &t0))) return rc;
    if ((rc = hc.get(hc.tev, 1, cudaEventDefault, &t1))) return rc;
    SCL_CUDA_TRY(cudaEventRecord(t0, stream));
  }
  // ONE tensor launch over all queries (one_chunk layout), signalling the completion of every query group.  The
  // counters are zeroed on this stream first; a consumer stream must not look at them before that (ev_armed).
  SCL_CUDA_TRY(cudaMemsetAsync(w.group_done, 0, size_t(kMaxGroups) * sizeof(unsigned int), stream));
  if (!hc.armed) SCL_CUDA_TRY(cudaEventCreateWithFlags(&hc.armed, cudaEventDisableTiming));
  SCL_CUDA_TRY(cudaEventRecord(hc.armed, stream));
  rc = launch_tensor(w, 0, w.qh, w.qmul, w.q_thr, Q, R, sv.Dp, sv.rn, sv.dbh, nullptr, false, stream, true);
  if (rc) return rc;
  if (timing) SCL_CUDA_TRY(cudaEventRecord(t1, stream));
  return SCL_OK;
}

extern "C" int scl_knn_query_begin_group(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q,
                                         int k, int group, float* ub, void* workspace, size_t workspace_bytes,
                                         scl_stream_t stream_) {
  if (!ub) return SCL_ERR_BAD_ARG;
  QueryWs w;
  int rc = two_phase_args(db, shadow, R, D, queries, Q, k, workspace, workspace_bytes, &w);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const ShardView sv = shard_view(shadow, R, D);
  const Groups g = group_layout(Q, sv.Dp);
  if (group >= g.n) return SCL_ERR_BAD_ARG;
  int mb, nt, NR, tpr, gm;
  knn_tc_tiling(Q, R, sv.Dp, &mb, &nt, &NR, &tpr, &gm);
  // wait, on THIS stream, until the launch of this thread has armed the counters and the tensor kernel has signalled
  // every work item of the group(s)
  HelperCtx& hc = t_helper[device_slot()];
  if (!hc.armed) return SCL_ERR_BAD_ARG;                   // no scl_knn_query_launch on this thread
  SCL_CUDA_TRY(cudaStreamWaitEvent(stream, hc.armed, 0));
  const int mu = g.unit_q == 256 ? (mb + 1) / 2 : mb;
  for (int gg = 0; gg < g.n; ++gg) {
    if (group >= 0 && gg != group) continue;
    const int units = std::min(gm, mu - gg * gm);
    if ((rc = stream_wait_geq(stream, w.group_done + gg, unsigned(units) * unsigned(NR) * unsigned(g.arrivals_per_item)))) return rc;
  }
  int q0, nq;
  group_range(g, Q, group, &q0, &nq);
  knn_cand_merge_kernel<<<nq, 256, 0, stream>>>(w.cand_s[0] + size_t(q0) * NR * kCandCap, w.cand_i[0] + size_t(q0) * NR * kCandCap,
                                                w.cand_cnt[0] + size_t(q0) * NR, w.q_thr + q0, NR, k, w.qn2 + q0, w.qexp + q0, sv.h,
                                                w.sel_idx + size_t(q0) * kKeep, w.sel_T + q0, w.sel_n + q0,
                                                w.sel_score + size_t(q0) * kKeep, ub);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

extern "C" int scl_knn_query_begin(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                                   float* ub, void* workspace, size_t workspace_bytes, scl_stream_t stream) {
  int rc = scl_knn_query_launch(db, shadow, R, D, queries, Q, k, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return scl_knn_query_begin_group(db, shadow, R, D, queries, Q, k, -1, ub, workspace, workspace_bytes, stream);
}

extern "C" int scl_knn_bound_reduce(const float* ub_all, int G, int Q, int k, float* bound, scl_stream_t stream) {
  if (!ub_all || !bound || G < 1 || Q < 1 || k < 1) return SCL_ERR_BAD_ARG;
  int rc = check_device();
  if (rc) return rc;
  knn_bound_reduce_kernel<<<(Q + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(ub_all, G, Q, k, bound);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

extern "C" int scl_knn_query_end_group(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                                       int64_t idx_offset, int group, const float* bound, double* dist, int64_t* idx,
                                       int32_t* stats, void* workspace, size_t workspace_bytes, scl_stream_t stream_) {
  if (!bound || !dist || !idx) return SCL_ERR_BAD_ARG;
  QueryWs w;
  int rc = two_phase_args(db, shadow, R, D, queries, Q, k, workspace, workspace_bytes, &w);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const ShardView sv = shard_view(shadow, R, D);
  const Groups g = group_layout(Q, sv.Dp);
  if (group >= g.n) return SCL_ERR_BAD_ARG;
  int q0, nq;
  group_range(g, Q, group, &q0, &nq);
  // everything below indexes the per-query arrays of the workspace with GLOBAL query numbers (the refused-query lists
  // hold global numbers); the caller's bound / dist / idx hold the group's rows only, so their bases are shifted back
  double* dist0 = dist - size_t(q0) * k;
  int64_t* idx0 = idx - size_t(q0) * k;
  SCL_CUDA_TRY(cudaMemsetAsync(w.stats, 0, 8 * sizeof(int), stream));
  knn_apply_cutoff_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(bound, w.sel_score + size_t(q0) * kKeep, w.sel_T + q0, w.sel_n + q0,
                                                            w.qn2 + q0, w.qexp + q0, sv.h, nq, w.sel_idx + size_t(q0) * kKeep,
                                                            w.cut_mode + q0, w.stats);
  SCL_LAUNCH_CHECK();
  knn_rescore_kernel<<<nq, 256, 0, stream>>>(db, queries + size_t(q0) * D, D, w.sel_idx + size_t(q0) * kKeep,
                                             w.d2 + size_t(q0) * kKeep);
  SCL_LAUNCH_CHECK();
  knn_finalize_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(w.sel_idx + size_t(q0) * kKeep, w.d2 + size_t(q0) * kKeep, w.sel_T + q0,
                                                        w.sel_n + q0, w.qn2 + q0, w.qexp + q0, sv.h, nq, k, idx_offset, 0, q0, dist,
                                                        reinterpret_cast<long long*>(idx), w.kth_d2 + q0, w.flag_list, w.stats,
                                                        w.cut_mode + q0);
  SCL_LAUNCH_CHECK();
  int hs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  SCL_CUDA_TRY(cudaMemcpyAsync(hs, w.stats, sizeof(hs), cudaMemcpyDeviceToHost, stream));
  SCL_CUDA_TRY(cudaStreamSynchronize(stream));
  const bool last = group < 0 || group == g.n - 1;
  if (last && g_knn_timing.load(std::memory_order_relaxed) != 0) {
    HelperCtx& hc = t_helper[device_slot()];
    if (hc.tev.size() >= 2) {
      SCL_CUDA_TRY(cudaEventSynchronize(hc.tev[1]));      // the tensor kernel may still be retiring on ITS stream
      float ms = 0.0f;
      SCL_CUDA_TRY(cudaEventElapsedTime(&ms, hc.tev[0], hc.tev[1]));
      std::lock_guard<std::mutex> lk(g_knn_timing_mu);
      g_knn_tc_ms_sum += double(ms);
      g_knn_tc_calls += 1;
    }
  }
  return resolve_refused(hs, db, R, D, queries, nq, k, idx_offset, 2, dist0, idx0, stats, w, sv.h, sv.rn, sv.dbh, sv.Dp, 1, stream);
}

extern "C" int scl_knn_query_end(const float* db, const void* shadow, int64_t R, int D, const float* queries, int Q, int k,
                                 int64_t idx_offset, const float* bound, double* dist, int64_t* idx, int32_t* stats,
                                 void* workspace, size_t workspace_bytes, scl_stream_t stream) {
  return scl_knn_query_end_group(db, shadow, R, D, queries, Q, k, idx_offset, -1, bound, dist, idx, stats, workspace,
                                 workspace_bytes, stream);
}

extern "C" int scl_knn_timing(int enable, double* tensor_pass_ms_sum, int* tensor_pass_calls) {
  std::lock_guard<std::mutex> lk(g_knn_timing_mu);
  if (tensor_pass_ms_sum) *tensor_pass_ms_sum = g_knn_tc_ms_sum;
  if (tensor_pass_calls) *tensor_pass_calls = g_knn_tc_calls;
  if (enable >= 0) {
    g_knn_timing.store(enable);
    g_knn_tc_ms_sum = 0.0;
    g_knn_tc_calls = 0;
  }
  return SCL_OK;
}

extern "C" int scl_topk_merge(const double* d_all, const int64_t* i_all, int G, int Q, int k, int64_t shard_stride,
                              double* d, int64_t* i, scl_stream_t stream) {
  if (!d_all || !i_all || !d || !i || G < 1 || Q < 1 || k < 1) return SCL_ERR_BAD_ARG;
  if (shard_stride == 0) shard_stride = (int64_t)Q * k;
  if (shard_stride < (int64_t)Q * k) return SCL_ERR_BAD_SHAPE;
  int rc = check_device();
  if (rc) return rc;
  topk_merge_kernel<<<(Q + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_all, reinterpret_cast<const long long*>(i_all), G, Q, k, (long long)shard_stride, d, reinterpret_cast<long long*>(i));
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

extern "C" int scl_geo_topn(const double* query_xy, const double* ref_xy, int Q, int64_t R, const int64_t* top_i, int k,
                            double* top_g_dists, int64_t* gt_i, double* gt_g_dist, scl_stream_t stream) {
  if (!query_xy || !ref_xy || !top_i || !top_g_dists || !gt_i || !gt_g_dist || Q < 1 || R < 1 || k < 1)
    return SCL_ERR_BAD_ARG;
  int rc = check_device();
  if (rc) return rc;
  geo_topn_kernel<<<Q, 256, 0, static_cast<cudaStream_t>(stream)>>>(query_xy, ref_xy, Q, R,
                                                                   reinterpret_cast<const long long*>(top_i), k, top_g_dists,
                                                                   reinterpret_cast<long long*>(gt_i), gt_g_dist);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}

extern "C" int scl_recall_curves(const double* top_g_dists, int Q, int k, const double* thresholds, int n_thresholds,
                                 double* curves, scl_stream_t stream_) {
  if (!top_g_dists || !thresholds || !curves || Q < 1 || k < 1 || n_thresholds < 1) return SCL_ERR_BAD_ARG;
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n = k * n_thresholds;
  recall_zero_kernel<<<(n + 255) / 256, 256, 0, stream>>>(curves, n);
  SCL_LAUNCH_CHECK();
  recall_count_kernel<<<(Q + 255) / 256, 256, 0, stream>>>(top_g_dists, Q, k, thresholds, n_thresholds, curves);
  SCL_LAUNCH_CHECK();
  recall_scale_kernel<<<(n + 255) / 256, 256, 0, stream>>>(curves, n, Q);
  SCL_LAUNCH_CHECK();
  return SCL_OK;
}
