// layernorm.cu — rms_norm and fused_add_rms_norm.
//
// Replaces csrc/layernorm_kernels.cu:22-46 (rms_norm_kernel), :187-273 (fused_add_rms_norm
// kernels) and the launchers :277-352.
//
// Reference numerics (kept):
//   rms_norm:            out = ((T)(x * rsqrt(mean(x^2) + eps))) * w      rounds to T BEFORE * w
//   fused_add_rms_norm:  z = x + r  (in T; the packed __hadd2 of the 128-bit path and the
//                        float-add-then-round of the generic path give identical bits for
//                        f16/bf16), residual = z, x = ((T)(z * s)) * w
//   the sum of squares is accumulated in fp32; only its association order is ours.
//
// B200 shape: HBM/latency bound, T*H*s*{2,4} bytes.  One CTA per token keeps the whole row in
// registers between the reduction and the scaling pass (one HBM read of x and r instead of the
// reference's two passes over global memory), 128-bit accesses, and a register-resident fast
// path for rows up to 8 vectors per thread at 256 threads (H <= 16384 for 2-byte types).
#include "common.cuh"

namespace vltr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// block-wide sum, result broadcast to every thread; red must hold 32 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = lane < nw ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// ---- vector path: row held in registers -------------------------------------------------------
// VPT = 16-byte vectors per thread (compile-time upper bound), FUSED = add residual first
template <typename T, int VPT, bool FUSED>
__global__ void __launch_bounds__(256)
    rms_vec_kernel(T* __restrict__ out, T* __restrict__ input, T* __restrict__ residual,
                   const T* __restrict__ weight, float epsilon, int hidden_size) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[32];
  const int nvec = hidden_size / N;
  const int64_t row = (int64_t)blockIdx.x * hidden_size;
  Vec16<T> x[VPT], w[VPT];
  float ss = 0.f;
  // the norm weight is a model constant: issued first so it is in flight together with the row
  // (and, where ptxas keeps it ahead of the wait, while the previous kernel is still draining)
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * blockDim.x;
    if (v < nvec) *reinterpret_cast<uint4*>(&w[k]) = ld_nc_16(weight + (int64_t)v * N);
  }
  pdl_wait();
  pdl_trigger();
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * blockDim.x;
    if (v < nvec) *reinterpret_cast<uint4*>(&x[k]) = ld_nc_16(input + row + (int64_t)v * N);
  }
  if constexpr (FUSED) {
    Vec16<T> r[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * blockDim.x;
      if (v < nvec)
        *reinterpret_cast<uint4*>(&r[k]) =
            *reinterpret_cast<const uint4*>(residual + row + (int64_t)v * N);
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * blockDim.x;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < N; ++j)
          x[k].v[j] = ST<T>::from_f(ST<T>::to_f(x[k].v[j]) + ST<T>::to_f(r[k].v[j]));
        st_16(residual + row + (int64_t)v * N, *reinterpret_cast<uint4*>(&x[k]));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * blockDim.x;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float f = ST<T>::to_f(x[k].v[j]);
        ss += f * f;
      }
    }
  }
  ss = block_sum(ss, red);
  const float s = rsqrtf(ss / hidden_size + epsilon);
  T* dst = FUSED ? input : out;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * blockDim.x;
    if (v < nvec) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float y = rt<T>(ST<T>::to_f(x[k].v[j]) * s);
        o.v[j] = ST<T>::from_f(y * ST<T>::to_f(w[k].v[j]));
      }
      st_16(dst + row + (int64_t)v * N, *reinterpret_cast<uint4*>(&o));
    }
  }
}

// ---- generic path: any hidden size / alignment (the reference's odd sizes 769, 5125, 8199) ----
template <typename T, bool FUSED>
__global__ void __launch_bounds__(1024)
    rms_generic_kernel(T* __restrict__ out, T* __restrict__ input, T* __restrict__ residual,
                       const T* __restrict__ weight, float epsilon, int hidden_size) {
  __shared__ float red[32];
  const int64_t row = (int64_t)blockIdx.x * hidden_size;
  float ss = 0.f;
  for (int i = threadIdx.x; i < hidden_size; i += blockDim.x) {
    float f = ST<T>::to_f(input[row + i]);
    if constexpr (FUSED) {
      const T z = ST<T>::from_f(f + ST<T>::to_f(residual[row + i]));
      residual[row + i] = z;
      f = ST<T>::to_f(z);
    }
    ss += f * f;
  }
  ss = block_sum(ss, red);
  const float s = rsqrtf(ss / hidden_size + epsilon);
  const T* src = FUSED ? residual : input;
  T* dst = FUSED ? input : out;
  for (int i = threadIdx.x; i < hidden_size; i += blockDim.x) {
    const float y = rt<T>(ST<T>::to_f(src[row + i]) * s);
    dst[row + i] = ST<T>::from_f(y * ST<T>::to_f(weight[i]));
  }
}

template <typename T, bool FUSED>
int launch_rms(void* out, void* input, void* residual, const void* weight, float eps,
               int num_tokens, int hidden, cudaStream_t stream) {
  constexpr int N = Vec16<T>::N;
  if (num_tokens == 0 || hidden == 0) return VLTR_OK;
  const bool aligned = hidden % N == 0 && (uintptr_t)input % 16 == 0 &&
                       (uintptr_t)weight % 16 == 0 && (FUSED ? (uintptr_t)residual % 16 == 0
                                                             : (uintptr_t)out % 16 == 0);
  const int nvec = hidden / N;
  T* o = (T*)out;
  T* in = (T*)input;
  T* res = (T*)residual;
  const T* w = (const T*)weight;
  if (aligned && nvec <= 8 * 256) {
    // threads: enough to give each thread >=1 vector, rounded to a warp multiple, max 256
    int threads = nvec >= 256 ? 256 : ((nvec + 31) / 32) * 32;
    const int vpt = (nvec + threads - 1) / threads;
    if (vpt <= 1)
      launch(rms_vec_kernel<T, 1, FUSED>, dim3(num_tokens), dim3(threads), 0, stream, o, in, res, w, eps, hidden);
    else if (vpt <= 2)
      launch(rms_vec_kernel<T, 2, FUSED>, dim3(num_tokens), dim3(threads), 0, stream, o, in, res, w, eps, hidden);
    else if (vpt <= 4)
      launch(rms_vec_kernel<T, 4, FUSED>, dim3(num_tokens), dim3(threads), 0, stream, o, in, res, w, eps, hidden);
    else
      launch(rms_vec_kernel<T, 8, FUSED>, dim3(num_tokens), dim3(threads), 0, stream, o, in, res, w, eps, hidden);
  } else {
    int threads = hidden < 1024 ? ((hidden + 31) / 32) * 32 : 1024;
    rms_generic_kernel<T, FUSED><<<num_tokens, threads, 0, stream>>>(o, in, res, w, eps, hidden);
  }
  VLTR_LAUNCH_CHECK(FUSED ? "fused_add_rms_norm" : "rms_norm");
  return VLTR_OK;
}

}  // namespace vltr

extern "C" int vltr_rms_norm(void* out, const void* input, const void* weight, float epsilon,
                             int num_tokens, int hidden_size, int dtype, void* stream) {
  using namespace vltr;
  VLTR_CHECK(num_tokens >= 0 && hidden_size >= 0, VLTR_ERR_INVALID, "rms_norm: negative size");
  VLTR_DISPATCH_FLOAT(dtype, "rms_norm", {
    return launch_rms<scalar_t, false>(out, const_cast<void*>(input), nullptr, weight, epsilon,
                                       num_tokens, hidden_size, (cudaStream_t)stream);
  });
  return VLTR_OK;
}

extern "C" int vltr_fused_add_rms_norm(void* input, void* residual, const void* weight,
                                       float epsilon, int num_tokens, int hidden_size, int dtype,
                                       void* stream) {
  using namespace vltr;
  VLTR_CHECK(num_tokens >= 0 && hidden_size >= 0, VLTR_ERR_INVALID,
             "fused_add_rms_norm: negative size");
  VLTR_DISPATCH_FLOAT(dtype, "fused_add_rms_norm", {
    return launch_rms<scalar_t, true>(nullptr, input, residual, weight, epsilon, num_tokens,
                                      hidden_size, (cudaStream_t)stream);
  });
  return VLTR_OK;
}
