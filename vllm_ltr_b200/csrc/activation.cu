// activation.cu — silu_and_mul / gelu_and_mul / gelu_tanh_and_mul / gelu_new / gelu_fast.
//
// Replaces csrc/activation_kernels.cu:13-161 of the reference.  Numerics follow the reference
// expression by expression ("in T" means float op + round-to-nearest back to T, which is what
// c10::Half / c10::BFloat16 operators do), so f16/bf16 results are bit-identical to the
// reference CUDA build: the reference test demands atol = rtol = 0 for silu
// (tests/kernels/test_activation.py:49).
//
// B200 shape: pure streaming, 3*s bytes per output element.  The reference runs one CTA per
// token with scalar 2-byte accesses; here the [tokens, d] problem is flattened into 16-byte
// vectors spread over the whole grid (one vector of x, one of y, one store per thread), so a
// decode batch of 256 tokens fills all 148 SMs instead of 256 short CTAs.
#include "common.cuh"

namespace vltr {

enum { ACT_SILU = 0, ACT_GELU = 1, ACT_GELU_TANH = 2, ACT_GELU_NEW = 3, ACT_GELU_FAST = 4 };

// gated activations return the value of ACT_FN(x) as the reference's T, held in float
template <typename T, int ACT>
__device__ __forceinline__ float act_fn(T xv) {
  const float f = ST<T>::to_f(xv);
  if constexpr (ACT == ACT_SILU) {
    // (T)(((float)x) / (1.0f + expf((float)-x)))          activation_kernels.cu:27-30
    return rt<T>(f / (1.0f + expf(-f)));
  } else if constexpr (ACT == ACT_GELU) {
    // (T)(f * 0.5f * (1.0f + ::erf(f * ALPHA)))            activation_kernels.cu:33-40
    constexpr float ALPHA = 0.70710678118654752440f;
    return rt<T>(f * 0.5f * (1.0f + erff(f * ALPHA)));
  } else if constexpr (ACT == ACT_GELU_TANH) {
    // activation_kernels.cu:43-53
    constexpr float BETA = (float)(1.41421356237309504880 * 1.12837916709551257390 * 0.5);
    constexpr float KAPPA = 0.044715f;
    float x_cube = f * f * f;
    float inner = BETA * (f + KAPPA * x_cube);
    return rt<T>(0.5f * f * (1.0f + tanhf(inner)));
  } else if constexpr (ACT == ACT_GELU_NEW) {
    // activation_kernels.cu:132-137: x*x*x, the inner sum and the products are all in T
    const float x2 = rt<T>(f * f);
    const float x3 = rt<T>(x2 * f);
    const float a = rt<T>(0.044715f * x3);
    const float b = rt<T>(f + a);
    const float c = rt<T>(0.79788456f * b);
    const float t = rt<T>(tanhf(c));
    const float h = rt<T>(rt<T>(0.5f) * f);
    const float u = rt<T>(1.0f + t);
    return rt<T>(h * u);
  } else {
    // activation_kernels.cu:139-144
    const float a = rt<T>(f * 0.79788456f);
    const float b = rt<T>(0.044715f * f);
    const float c = rt<T>(b * f);
    const float d = rt<T>(1.0f + c);
    const float e = rt<T>(a * d);
    const float t = rt<T>(tanhf(e));
    const float h = rt<T>(rt<T>(0.5f) * f);
    const float u = rt<T>(1.0f + t);
    return rt<T>(h * u);
  }
}

// vectorised: one 16-byte vector of out per thread
template <typename T, int ACT, bool GATED>
__global__ void __launch_bounds__(256) act_vec_kernel(T* __restrict__ out,
                                                      const T* __restrict__ input,
                                                      int64_t num_vec, int vec_per_row, int d) {
  constexpr int N = Vec16<T>::N;
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_vec) return;
  const int64_t row = i / vec_per_row;
  const int col = (int)(i - row * vec_per_row) * N;
  const int64_t in_row = GATED ? row * 2 * (int64_t)d : row * (int64_t)d;
  Vec16<T> x, y, o;
  *reinterpret_cast<uint4*>(&x) = ld_nc_16(input + in_row + col);
  if constexpr (GATED) *reinterpret_cast<uint4*>(&y) = ld_nc_16(input + in_row + d + col);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float a = act_fn<T, ACT>(x.v[j]);
    if constexpr (GATED) a = a * ST<T>::to_f(y.v[j]);  // ACT_FN(x) * y in T
    o.v[j] = ST<T>::from_f(a);
  }
  st_16(out + row * (int64_t)d + col, *reinterpret_cast<uint4*>(&o));
}

// scalar path for rows that are not 16-byte tileable
template <typename T, int ACT, bool GATED>
__global__ void __launch_bounds__(256) act_scalar_kernel(T* __restrict__ out,
                                                         const T* __restrict__ input,
                                                         int64_t total, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t row = i / d;
  const int col = (int)(i - row * d);
  const int64_t in_row = GATED ? row * 2 * (int64_t)d : row * (int64_t)d;
  float a = act_fn<T, ACT>(input[in_row + col]);
  if constexpr (GATED) a = a * ST<T>::to_f(input[in_row + d + col]);
  out[i] = ST<T>::from_f(a);
}

template <typename T, int ACT, bool GATED>
int launch_act(void* out, const void* input, int64_t num_tokens, int d, cudaStream_t stream) {
  constexpr int N = Vec16<T>::N;
  if (num_tokens == 0 || d == 0) return VLTR_OK;
  const bool aligned = (d % N == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)input % 16 == 0);
  if (aligned) {
    const int vec_per_row = d / N;
    const int64_t num_vec = num_tokens * vec_per_row;
    const int64_t grid = (num_vec + 255) / 256;
    launch(act_vec_kernel<T, ACT, GATED>, dim3((unsigned)grid), dim3(256), 0, stream, (T*)out,
           (const T*)input, num_vec, vec_per_row, d);
  } else {
    const int64_t total = num_tokens * d;
    const int64_t grid = (total + 255) / 256;
    act_scalar_kernel<T, ACT, GATED><<<(unsigned)grid, 256, 0, stream>>>((T*)out, (const T*)input,
                                                                         total, d);
  }
  VLTR_LAUNCH_CHECK("activation");
  return VLTR_OK;
}

}  // namespace vltr

extern "C" int vltr_activation(void* out, const void* input, int64_t num_tokens, int d, int act,
                               int dtype, void* stream) {
  using namespace vltr;
  cudaStream_t s = (cudaStream_t)stream;
  VLTR_CHECK(num_tokens >= 0 && d >= 0, VLTR_ERR_INVALID, "activation: negative size");
  VLTR_DISPATCH_FLOAT(dtype, "activation", {
    switch (act) {
      case ACT_SILU: return launch_act<scalar_t, ACT_SILU, true>(out, input, num_tokens, d, s);
      case ACT_GELU: return launch_act<scalar_t, ACT_GELU, true>(out, input, num_tokens, d, s);
      case ACT_GELU_TANH:
        return launch_act<scalar_t, ACT_GELU_TANH, true>(out, input, num_tokens, d, s);
      case ACT_GELU_NEW:
        return launch_act<scalar_t, ACT_GELU_NEW, false>(out, input, num_tokens, d, s);
      case ACT_GELU_FAST:
        return launch_act<scalar_t, ACT_GELU_FAST, false>(out, input, num_tokens, d, s);
      default: return fail(VLTR_ERR_INVALID, "activation: unknown act %d", act);
    }
  });
  return VLTR_OK;
}
