// Error plumbing, device facts and the measurement probe of libpybot_b200.
#include <stdarg.h>
#include <string.h>

#include "pbk_common.cuh"

namespace pbk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return PBK_ECUDA;
}

static DeviceInfo g_info[64];
static bool g_info_ok[64];

int current_device_info(const DeviceInfo** out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  PBK_REQUIRE(dev >= 0 && dev < 64, PBK_EINVAL, "device index %d out of range", dev);
  if (!g_info_ok[dev]) {
    DeviceInfo d;
    d.device = dev;
    PBK_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    PBK_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    PBK_CUDA(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    PBK_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    PBK_CUDA(cudaDeviceGetAttribute(&d.l2_bytes, cudaDevAttrL2CacheSize, dev));
    PBK_REQUIRE(d.cc_major == 10, PBK_ECUDA,
                "libpybot_b200 is built for sm_100a only; device %d is sm_%d%d",
                dev, d.cc_major, d.cc_minor);
    g_info[dev] = d;
    g_info_ok[dev] = true;
  }
  *out = &g_info[dev];
  return PBK_OK;
}

}  // namespace pbk

extern "C" {

int pbk_version(void) { return PBK_VERSION; }

const char* pbk_last_error_string(void) { return pbk::g_err; }

int pbk_device_props(int device, int64_t props[8]) {
  PBK_REQUIRE(props != nullptr, PBK_EINVAL, "props is NULL");
  int v[8];
  const cudaDeviceAttr a[8] = {
      cudaDevAttrMultiProcessorCount,      cudaDevAttrComputeCapabilityMajor,
      cudaDevAttrComputeCapabilityMinor,   cudaDevAttrMaxSharedMemoryPerBlockOptin,
      cudaDevAttrL2CacheSize,              cudaDevAttrClockRate,
      cudaDevAttrMemoryClockRate,          cudaDevAttrGlobalMemoryBusWidth};
  for (int i = 0; i < 8; ++i) {
    PBK_CUDA(cudaDeviceGetAttribute(&v[i], a[i], device));
    props[i] = v[i];
  }
  return PBK_OK;
}

}  // extern "C"
