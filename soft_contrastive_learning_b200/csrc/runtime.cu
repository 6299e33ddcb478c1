// runtime.cu -- status strings, device check, thread-local error detail.
#include <atomic>

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
