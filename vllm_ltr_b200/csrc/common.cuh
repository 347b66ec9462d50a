// common.cuh — shared helpers for libvllm_ltr_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/vllm_ltr_b200.h"

namespace vltr {

// ---- error reporting (thread-local message, reference: TORCH_CHECK -> RuntimeError) ----------
char* err_buf();
int fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define VLTR_CHECK(cond, code, ...)                 \
  do {                                              \
    if (!(cond)) return ::vltr::fail(code, __VA_ARGS__); \
  } while (0)

// call after every <<<>>>: counts the launch and converts a launch failure into an error code
#define VLTR_LAUNCH_CHECK(name)                                                        \
  do {                                                                                 \
    ::vltr::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
    cudaError_t e__ = cudaPeekAtLastError();                                           \
    if (e__ != cudaSuccess) {                                                          \
      cudaGetLastError();                                                              \
      return ::vltr::fail(VLTR_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__));     \
    }                                                                                  \
  } while (0)

#define VLTR_CUDA(call)                                                                \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return ::vltr::fail(VLTR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));    \
  } while (0)

// ---- launches --------------------------------------------------------------------------------
// Every hot-path kernel is launched with the programmatic-stream-serialization attribute
// (programmatic dependent launch, PDL): the grid may be scheduled while the previous kernel on
// the stream is still draining, and blocks in pdl_wait() until that kernel has completed and
// its writes are visible.  Kernels call pdl_wait() before their first read of data a
// predecessor may have written and pdl_trigger() right after it, so a chain of short kernels
// (one decoder layer = rms_norm, rotary, reshape_and_cache, attention, rms_norm, silu) overlaps
// each launch + prologue with the tail of the kernel before it.  Both instructions are no-ops
// when the predecessor is not PDL-aware (cuBLAS GEMMs of the engine) or the attribute is off
// (vltr_tune("pdl", 0) / VLTR_PDL=0).
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// development knobs (vltr_tune): 0 = library default
struct Tune {
  long long attn_warps, attn_stages, attn_ctas, attn_slots, attn_cap_tokens, attn_min_tokens;
  long long attn_trace;      // device pointer: per-warp globaltimer stamps (tools/attn_tune.py)
  long long opt_gemm;        // 1: one-tile-per-CTA predictor GEMM instead of the persistent one
  long long attn_split;      // paged attention two-launch (split) mode: 0 policy, 1 always, -1 never
  long long tp_multimem;     // -1: ignore multicast mappings (peer loads / stores even when NVLS is there)
  long long tp_trace;        // device pointer: 8 globaltimer stamps of the last exchange (tools/tp_check.py)
};
extern Tune g_tune;

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- scalar type traits -----------------------------------------------------------------------
// The reference computes "in T": every c10::Half / c10::BFloat16 operator converts to float,
// applies the op and rounds back (RN).  to_f / from_f are those two conversions.
template <typename T>
struct ST;
template <>
struct ST<float> {
  static constexpr int code = VLTR_F32;
  __device__ __forceinline__ static float to_f(float x) { return x; }
  __device__ __forceinline__ static float from_f(float x) { return x; }
};
template <>
struct ST<__half> {
  static constexpr int code = VLTR_F16;
  __device__ __forceinline__ static float to_f(__half x) { return __half2float(x); }
  __device__ __forceinline__ static __half from_f(float x) { return __float2half_rn(x); }
};
template <>
struct ST<__nv_bfloat16> {
  static constexpr int code = VLTR_BF16;
  __device__ __forceinline__ static float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
  __device__ __forceinline__ static __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
};

// round-trip through T: the value a reference expression "(scalar_t)(expr)" would hold, as float
template <typename T>
__device__ __forceinline__ float rt(float x) {
  return ST<T>::to_f(ST<T>::from_f(x));
}

// 16-byte vector of T
template <typename T>
struct alignas(16) Vec16 {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_16(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// ---- fp8 e5m2 (unscaled), csrc/quantization/fp8_e5m2_kvcache/quant_utils.cuh -----------------
// to fp8: saturate-to-finite RN conversion from the activation type (float / half / bf16 each
// have their own direct intrinsic in the reference, :164-192); from fp8: via half (:24-30),
// bf16 goes fp8 -> half -> float -> bf16 (:71-82).
template <typename T>
__device__ __forceinline__ uint8_t to_fp8(T x);
template <>
__device__ __forceinline__ uint8_t to_fp8<float>(float x) {
  return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E5M2);
}
template <>
__device__ __forceinline__ uint8_t to_fp8<__half>(__half x) {
  return (uint8_t)__nv_cvt_halfraw_to_fp8(static_cast<__half_raw>(x), __NV_SATFINITE, __NV_E5M2);
}
template <>
__device__ __forceinline__ uint8_t to_fp8<__nv_bfloat16>(__nv_bfloat16 x) {
  return (uint8_t)__nv_cvt_bfloat16raw_to_fp8(static_cast<__nv_bfloat16_raw>(x), __NV_SATFINITE,
                                              __NV_E5M2);
}
// e5m2 is the upper byte of an IEEE half: the reference's __nv_cvt_fp8_to_halfraw(a, __NV_E5M2)
// is exactly (a << 8).
__device__ __forceinline__ __half fp8_to_half(uint8_t a) {
  return __ushort_as_half((unsigned short)((unsigned short)a << 8));
}
template <typename T>
__device__ __forceinline__ T from_fp8(uint8_t a);
template <>
__device__ __forceinline__ float from_fp8<float>(uint8_t a) {
  return __half2float(fp8_to_half(a));
}
template <>
__device__ __forceinline__ __half from_fp8<__half>(uint8_t a) {
  return fp8_to_half(a);
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_fp8<__nv_bfloat16>(uint8_t a) {
  return __float2bfloat16(__half2float(fp8_to_half(a)));
}

// ---- helpers shared by pos_encoding.cu, cache.cu and rope_cache.cu --------------------------------
// one rotary pair "in T" (pos_encoding_kernels.cu:34-38: every * and +/- rounds to T)
template <typename T>
__device__ __forceinline__ void rot_pair(T& x, T& y, T c, T s) {
  const float fx = ST<T>::to_f(x), fy = ST<T>::to_f(y), fc = ST<T>::to_f(c), fs = ST<T>::to_f(s);
  if constexpr (sizeof(T) == 4) {
    x = fx * fc - fy * fs;
    y = fy * fc + fx * fs;
  } else {
    x = ST<T>::from_f(rt<T>(fx * fc) - rt<T>(fy * fs));
    y = ST<T>::from_f(rt<T>(fy * fc) + rt<T>(fx * fs));
  }
}

// element on its way into the KV cache: as is, or e5m2 (saturate to finite)
template <typename T, typename CT, bool FP8>
__device__ __forceinline__ CT cache_cvt(T v) {
  if constexpr (FP8)
    return to_fp8<T>(v);
  else
    return v;
}

}  // namespace vltr

// dtype dispatch over the three activation types
#define VLTR_DISPATCH_FLOAT(dtype, NAME, ...)                                   \
  switch (dtype) {                                                              \
    case VLTR_F32: {                                                            \
      using scalar_t = float;                                                   \
      __VA_ARGS__;                                                              \
      break;                                                                    \
    }                                                                           \
    case VLTR_F16: {                                                            \
      using scalar_t = __half;                                                  \
      __VA_ARGS__;                                                              \
      break;                                                                    \
    }                                                                           \
    case VLTR_BF16: {                                                           \
      using scalar_t = __nv_bfloat16;                                           \
      __VA_ARGS__;                                                              \
      break;                                                                    \
    }                                                                           \
    default:                                                                    \
      return ::vltr::fail(VLTR_ERR_UNSUPPORTED, "%s: Unsupported data type: %d", NAME, dtype); \
  }
