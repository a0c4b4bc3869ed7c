// Library-level plumbing of libkfb: error text, device queries.
#include <stdarg.h>

#include <mutex>

#include "kfb_common.cuh"

namespace kfb {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

}  // namespace kfb

extern "C" {

int kfb_version(void) { return KFB_VERSION; }

const char* kfb_last_error(void) { return kfb::g_error; }

void kfb_struct_sizes(int* layer, int* split, int* epilogue) {
  if (layer) *layer = (int)sizeof(kfb_layer);
  if (split) *split = (int)sizeof(kfb_split);
  if (epilogue) *epilogue = (int)sizeof(kfb_epilogue);
}

int kfb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    kfb::set_error("no CUDA device is visible: libkfb has no CPU path");
    return KFB_ERR_NO_DEVICE;
  }
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    kfb::set_error("cudaGetDeviceProperties failed");
    return KFB_ERR_CUDA;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (prop.major != 10) {
    kfb::set_error("device %d is sm_%d%d; libkfb is built for sm_100a only", dev, prop.major, prop.minor);
    return KFB_ERR_NO_DEVICE;
  }
  return KFB_OK;
}

}  // extern "C"
