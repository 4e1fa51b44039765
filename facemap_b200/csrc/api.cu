// Library-level pieces of the C-ABI: error text, launch counter, device check.
#include <stdarg.h>
#include "common.cuh"

namespace fm {

static thread_local char g_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace fm

extern "C" int fm_version(void) { return 100; }

extern "C" const char* fm_last_error(void) { return fm::g_error; }

extern "C" uint64_t fm_launch_count(void) { return fm::g_launches.load(std::memory_order_relaxed); }

extern "C" int fm_check_device(int dev) {
  using namespace fm;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible: facemap_b200 has no CPU fallback");
    return FM_ERR_NO_DEVICE;
  }
  FM_REQUIRE(dev >= 0 && dev < count, "device %d out of range (%d visible)", dev, count);
  cudaDeviceProp prop;
  FM_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library carries sm_100a code only", dev, prop.major, prop.minor);
    return FM_ERR_NO_DEVICE;
  }
  return FM_OK;
}
