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

// Page-lock caller memory in place so H2D copies from it are true DMA.  A refused registration
// (already registered, not lockable) is not an error of the pipeline: the sticky CUDA error is
// cleared and the caller falls back to staged copies.
extern "C" int fm_host_register(void* ptr, size_t bytes) {
  using namespace fm;
  FM_REQUIRE(ptr != nullptr && bytes > 0, "fm_host_register: bad args");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaHostRegister refused: %s", cudaGetErrorString(e));
    return FM_ERR_CUDA;
  }
  return FM_OK;
}

extern "C" int fm_host_unregister(void* ptr) {
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return FM_ERR_CUDA;
  }
  return FM_OK;
}
