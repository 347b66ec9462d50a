// runtime.cu — error channel, launch accounting and the cuda_utils pair
// (reference: csrc/cuda_utils_kernels.cu:1-35).
#include <cstdlib>
#include <string>

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

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v != nullptr && *v != 0) ? atoi(v) : dflt;
}
int g_pdl = env_int("VLTR_PDL", 1);
Tune g_tune{};

}  // namespace vltr

extern "C" {

const char* vltr_last_error(void) { return vltr::err_buf(); }

const char* vltr_build_info(void) { return "sm_100a;" __DATE__ " " __TIME__; }

uint64_t vltr_launch_count(void) { return vltr::g_launches.load(std::memory_order_relaxed); }

int vltr_tune(const char* key, long long value) {
  using vltr::g_tune;
  const std::string k(key ? key : "");
  if (k == "pdl") vltr::g_pdl = value != 0;
  else if (k == "attn_warps") g_tune.attn_warps = value;
  else if (k == "attn_stages") g_tune.attn_stages = value;
  else if (k == "attn_ctas") g_tune.attn_ctas = value;
  else if (k == "attn_slots") g_tune.attn_slots = value;
  else if (k == "attn_cap_tokens") g_tune.attn_cap_tokens = value;
  else if (k == "attn_min_tokens") g_tune.attn_min_tokens = value;
  else if (k == "attn_trace") g_tune.attn_trace = value;
  else if (k == "opt_gemm") g_tune.opt_gemm = value;
  else if (k == "attn_split") g_tune.attn_split = value;
  else if (k == "tp_multimem") g_tune.tp_multimem = value;
  else if (k == "tp_trace") g_tune.tp_trace = value;
  else return vltr::fail(VLTR_ERR_INVALID, "vltr_tune: unknown key '%s'", k.c_str());
  return VLTR_OK;
}

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
