// runtime.cu -- status strings, device check, thread-local error detail.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace scl {

static thread_local char g_err[512] = "";

void set_last_error(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

// Per-device capability cache (index = device ordinal).
static std::atomic<int> g_cc[64];       // 0 unknown, else major*10+minor
static std::atomic<int> g_sms[64];

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_last_error("cudaGetDevice", e);
    return SCL_ERR_CUDA;
  }
  if (dev < 0 || dev >= 64) return SCL_ERR_ARCH;
  int cc = g_cc[dev].load(std::memory_order_relaxed);
  if (cc == 0) {
    int major = 0, minor = 0, sms = 0;
    if ((e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) {
      set_last_error("cudaDeviceGetAttribute", e);
      return SCL_ERR_CUDA;
    }
    cc = major * 10 + minor;
    g_sms[dev].store(sms, std::memory_order_relaxed);
    g_cc[dev].store(cc, std::memory_order_relaxed);
  }
  return cc / 10 == 10 ? SCL_OK : SCL_ERR_ARCH;   // sm_100a code only runs on compute capability 10.x
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int ensure_dyn_smem(const void* func, size_t bytes, SmemAttrCache* cache) {
  if (bytes <= 48 * 1024) return SCL_OK;
  const int dev = device_slot();
  // the attribute only ever grows and setting it twice is harmless
  if (cache->bytes[dev].load(std::memory_order_relaxed) >= bytes) return SCL_OK;
  SCL_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
  cache->bytes[dev].store(bytes, std::memory_order_relaxed);
  return SCL_OK;
}

// ---------------------------------------------------------------------------------------------
// tuning knobs: environment read once at load, then scl_set_tuning only
// ---------------------------------------------------------------------------------------------
static const char* const kKnobNames[KNOB_COUNT] = {
    "SCL_WMS_STREAM", "SCL_WMS_STREAM_CFG", "SCL_WMS_CLUSTER", "SCL_WMS_CHUNKED", "SCL_TUPLE_CLUSTER",
    "SCL_KNN_TC_VARIANT", "SCL_KNN_SYNC", "SCL_KNN_SYNC_WINDOW", "SCL_KNN_SYNC_SUBS", "SCL_KNN_RANGES",
    "SCL_KNN_GROUP_M", "SCL_KNN_CHUNK_Q", "SCL_KNN_STAGE2", "SCL_GEMM_SIMT", "SCL_NV_FUSED"};
struct KnobTable {
  std::atomic<int> v[KNOB_COUNT];
  KnobTable() {
    for (int i = 0; i < KNOB_COUNT; ++i) {
      const char* e = getenv(kKnobNames[i]);
      v[i].store((e && *e) ? atoi(e) : kKnobUnset, std::memory_order_relaxed);
    }
  }
};
static KnobTable g_knobs;       // constructed when the shared library is loaded

int knob(Knob k) { return g_knobs.v[k].load(std::memory_order_relaxed); }

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (g_cc[dev].load(std::memory_order_relaxed) == 0) check_device();
  int s = g_sms[dev].load(std::memory_order_relaxed);
  return s > 0 ? s : 148;
}

}  // namespace scl

extern "C" int scl_version(void) { return 100; }

extern "C" const char* scl_strerror(int status) {
  switch (status) {
    case SCL_OK: return "ok";
    case SCL_ERR_BAD_ARG: return "bad argument (null pointer or unknown enum)";
    case SCL_ERR_BAD_SHAPE: return "unsupported or inconsistent shape";
    case SCL_ERR_ALIGN: return "pointer not 16-byte aligned";
    case SCL_ERR_WORKSPACE: return "workspace too small";
    case SCL_ERR_CUDA: return "CUDA call failed (see scl_last_error)";
    case SCL_ERR_ARCH: return "device is not compute capability 10.x (sm_100a build, no fallback)";
    case SCL_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

extern "C" const char* scl_last_error(void) { return scl::g_err; }

extern "C" int scl_device_ok(void) { return scl::check_device(); }

extern "C" int scl_set_tuning(const char* name, int value) {
  if (!name) return SCL_ERR_BAD_ARG;
  for (int i = 0; i < scl::KNOB_COUNT; ++i)
    if (strcmp(name, scl::kKnobNames[i]) == 0) {
      scl::g_knobs.v[i].store(value, std::memory_order_relaxed);
      return SCL_OK;
    }
  return SCL_ERR_BAD_ARG;
}
extern "C" int scl_get_tuning(const char* name, int* value) {
  if (!name || !value) return SCL_ERR_BAD_ARG;
  for (int i = 0; i < scl::KNOB_COUNT; ++i)
    if (strcmp(name, scl::kKnobNames[i]) == 0) {
      *value = scl::g_knobs.v[i].load(std::memory_order_relaxed);
      return SCL_OK;
    }
  return SCL_ERR_BAD_ARG;
}
