// runtime.cu — error channel, launch accounting and the cuda_utils pair
// (reference: csrc/cuda_utils_kernels.cu:1-35).
#include "common.cuh"

namespace vltr {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

std::atomic<uint64_t> g_launches{0};

}  // namespace vltr

extern "C" {

const char* vltr_last_error(void) { return vltr::err_buf(); }

const char* vltr_build_info(void) { return "sm_100a;" __DATE__ " " __TIME__; }

uint64_t vltr_launch_count(void) { return vltr::g_launches.load(std::memory_order_relaxed); }

int vltr_get_device_attribute(int attribute, int device_id) {
  int device = device_id, value = 0;
  if (device_id < 0) cudaGetDevice(&device);
  cudaDeviceGetAttribute(&value, static_cast<cudaDeviceAttr>(attribute), device);
  return value;
}

int vltr_get_max_shared_memory_per_block_device_attribute(int device_id) {
  return vltr_get_device_attribute((int)cudaDevAttrMaxSharedMemoryPerBlockOptin, device_id);
}

}  // extern "C"
