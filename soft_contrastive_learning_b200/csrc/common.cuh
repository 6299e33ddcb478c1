// common.cuh -- shared helpers for libscl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/scl_b200.h"

namespace scl {

void set_last_error(const char* what, cudaError_t e);
int check_device();  // SCL_OK or SCL_ERR_ARCH / SCL_ERR_CUDA
int num_sms();
int device_slot();   // ordinal of the current device, clamped to [0, kMaxDevices)
constexpr int kMaxDevices = 64;

// Process-wide tuning knobs (include/scl_b200.h: scl_set_tuning / scl_get_tuning).  Each knob is read from its SCL_*
// environment variable ONCE, when the library is loaded; afterwards only scl_set_tuning changes it.  No entry point calls
// getenv.
enum Knob {
  KNOB_WMS_STREAM = 0,      // 0 never / 1 always use the streaming wms kernel (unset: batches >= #SMs tuples)
  KNOB_WMS_STREAM_CFG,      // streaming kernel variant (wms_tuple_stream.cu)
  KNOB_WMS_CLUSTER,         // CTAs per tuple of the cluster kernels (1, 2, 4, 8)
  KNOB_WMS_CHUNKED,         // 1: skip the resident kernel
  KNOB_TUPLE_CLUSTER,       // CTAs per tuple of the triplet-family kernel
  KNOB_KNN_TC_VARIANT,      // 1 single CTA, 2 CTA pair (default), 3 CTA pair 256x512
  KNOB_KNN_SYNC, KNOB_KNN_SYNC_WINDOW, KNOB_KNN_SYNC_SUBS, KNOB_KNN_RANGES, KNOB_KNN_GROUP_M,
  KNOB_KNN_CHUNK_Q,         // queries per pipelined chunk of scl_knn_query (0: one chunk)
  KNOB_KNN_STAGE2,          // 0: uncertified queries go straight to the exact scan (skip the split-fp16 second stage)
  KNOB_GEMM_SIMT,           // 1: FP32 FFMA GEMM instead of tcgen05
  KNOB_NV_FUSED,            // 0: NetVLAD through the generic GEMM engine instead of the fused kernels
  KNOB_COUNT
};
constexpr int kKnobUnset = -2147483647 - 1;
int knob(Knob k);                                    // kKnobUnset when not set
static inline int knob_or(Knob k, int dflt) { const int v = knob(k); return v == kKnobUnset ? dflt : v; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: the cache of "already raised
// to X bytes" is keyed by the device ordinal.  Idempotent; a race between host threads is benign.
struct SmemAttrCache {
  std::atomic<size_t> bytes[kMaxDevices];
};
int ensure_dyn_smem(const void* func, size_t bytes, SmemAttrCache* cache);

#define SCL_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) {                                 \
      ::scl::set_last_error(#expr, _e);                      \
      return SCL_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

#define SCL_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) {                                 \
      ::scl::set_last_error("kernel launch", _e);            \
      return SCL_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (256-byte granules).
struct Carver {
  char* base;
  size_t off;
  size_t cap;
  Carver(void* p, size_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes) {}
  template <typename T>
  T* take(size_t n) {
    size_t b = round_up(n * sizeof(T), 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += b;
    return r;
  }
  bool ok() const { return off <= cap; }
};
static inline size_t carve_bytes(size_t n, size_t elem) { return round_up(n * elem, 256); }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// streaming 128-bit global load / store (read-once data: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
#endif

}  // namespace scl
