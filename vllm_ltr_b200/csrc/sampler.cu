// sampler.cu — next-token sampling from the logits rows of one engine step, in one kernel.
//
// Replaces, for the requests the throughput benchmarks issue (greedy, or random sampling with
// temperature; no penalties / top-k / top-p / min-p), the tensor program of
// vllm/model_executor/layers/sampler.py:
//   :67   logits.div_(temperatures)                       in T (the temperature is a T tensor)
//   :77   probs    = softmax(logits, fp32)                [R, V] fp32 written
//   :80   logprobs = log_softmax(logits, fp32)            [R, V] fp32 written
//   :395  greedy:   argmax(logprobs)                      first maximal index
//   :333-357 random: q = empty_like(probs).exponential_(); argmax(probs / q)
//   :528-544, :594-600  logprob and rank of the chosen token: logprobs[r, tok],
//                        (logprobs[r] > logprobs[r, tok]).sum() + 1
// About 25 bytes of HBM traffic per logit and a dozen launches there; here one CTA per row reads
// the row ONCE from HBM (2 or 4 bytes per logit) keeping an online (max, sum) pair and the running
// arg-max of the sampling key, then counts the rank in a second sweep that hits L2.
//
// Random sampling uses the same exponential-race construction as the reference:
//   argmax_i p_i / q_i,  q_i ~ Exp(1)   ==   argmax_i (x_i - log q_i)       (monotone transform)
// The noise is either supplied by the caller (`noise` [R, V] fp32 exponential samples - what the
// parity tests do, against the reference run on the same q) or drawn in the kernel from
// Philox4x32-10 keyed by (seed, row, column): q = -log(u), u in (0, 1].
//
// Bound: HBM, V*sizeof(T) bytes per row (+ 4*V when the caller supplies noise or asks for the
// full probs / logprobs matrices).
#include "common.cuh"

namespace vltr {
namespace sampler {

constexpr int kThreads = 512;

// ---- Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1) -----------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// uniform in (0, 1]: never 0, so -log(u) is finite
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 1.0f) * (1.0f / 16777216.0f); }

struct Acc {
  float m, s;        // online softmax: running max, sum of exp(x - m)
  float key;         // best sampling key so far
  int idx;           // its column (smallest column on ties)
};
__device__ __forceinline__ void acc_merge(Acc& a, const Acc& b) {
  const float m = fmaxf(a.m, b.m);
  // exp(-inf - -inf) guards: an empty side contributes nothing
  a.s = (a.m == -INFINITY ? 0.f : a.s * __expf(a.m - m)) + (b.m == -INFINITY ? 0.f : b.s * __expf(b.m - m));
  a.m = m;
  if (b.key > a.key || (b.key == a.key && b.idx < a.idx)) {
    a.key = b.key;
    a.idx = b.idx;
  }
}
__device__ __forceinline__ void acc_add(Acc& a, float x, float key, int col) {
  if (x > a.m) {
    a.s = a.s * __expf(a.m - x) + 1.0f;   // a.m == -inf: 0 * 0 + 1
    a.m = x;
  } else if (x != -INFINITY) {
    a.s += __expf(x - a.m);
  }
  if (key > a.key) {   // columns arrive in increasing order per thread: strict > keeps the first
    a.key = key;
    a.idx = col;
  }
}

template <typename T>
__device__ __forceinline__ float scaled(T v, float t, bool unit) {
  // logits.div_(temperature) in T: fp32 divide, rounded to T (sampler.py:67)
  const float f = ST<T>::to_f(v);
  return unit ? f : rt<T>(f / t);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    sample_kernel(const T* __restrict__ logits, int64_t row_stride, int vocab,
                  const float* __restrict__ temperatures, const uint8_t* __restrict__ random,
                  const float* __restrict__ noise, uint64_t seed, uint64_t offset,
                  int64_t* __restrict__ out_token, float* __restrict__ out_logprob,
                  int64_t* __restrict__ out_rank, float* __restrict__ out_probs,
                  float* __restrict__ out_logprobs) {
  constexpr int N = 16 / sizeof(T);
  __shared__ Acc red[kThreads / 32];
  __shared__ int cnt[kThreads / 32];
  __shared__ Acc fin;
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // temperatures arrive as the T tensor the reference builds (SamplingTensors, dtype = logits
  // dtype): the caller passes them already rounded to T, as floats
  const float t = temperatures != nullptr ? temperatures[row] : 1.0f;
  const bool unit = t == 1.0f;
  const bool rnd = random != nullptr && random[row] != 0;
  const T* x = logits + (int64_t)row * row_stride;
  const float* nz = noise != nullptr ? noise + (int64_t)row * vocab : nullptr;
  const uint2 pkey = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  pdl_wait();
  pdl_trigger();

  Acc a{-INFINITY, 0.f, -INFINITY, 0x7fffffff};
  const bool vec_ok = (((uintptr_t)x) % 16 == 0);
  const int nvec = vec_ok ? vocab / N : 0;
  // ---- sweep 1: HBM -> (max, sum, arg-max of the key) ------------------------------------------
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    Vec16<T> c;
    *reinterpret_cast<uint4*>(&c) = ld_nc_16(x + (int64_t)v * N);
    float q[N];
    if (rnd) {
      if (nz != nullptr) {
#pragma unroll
        for (int j = 0; j < N; ++j) q[j] = nz[v * N + j];
      } else {
#pragma unroll
        for (int j4 = 0; j4 < N; j4 += 4) {
          const uint32_t ctr = (uint32_t)((v * N + j4) >> 2);
          const uint4 r = philox4x32_10(make_uint4(ctr, (uint32_t)row, (uint32_t)offset,
                                                   (uint32_t)(offset >> 32)), pkey);
          q[j4] = -__logf(u01(r.x));
          q[j4 + 1] = -__logf(u01(r.y));
          q[j4 + 2] = -__logf(u01(r.z));
          q[j4 + 3] = -__logf(u01(r.w));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float f = scaled<T>(c.v[j], t, unit);
      const float key = rnd ? (f == -INFINITY ? -INFINITY : f - __logf(q[j])) : f;
      acc_add(a, f, key, v * N + j);
    }
  }
  for (int i = nvec * N + threadIdx.x; i < vocab; i += kThreads) {   // tail / unaligned rows
    const float f = scaled<T>(x[i], t, unit);
    float key = f;
    if (rnd && f != -INFINITY) {
      float q;
      if (nz != nullptr) {
        q = nz[i];
      } else {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)(i >> 2), (uint32_t)row, (uint32_t)offset,
                                                 (uint32_t)(offset >> 32)), pkey);
        const uint32_t w = (i & 3) == 0 ? r.x : ((i & 3) == 1 ? r.y : ((i & 3) == 2 ? r.z : r.w));
        q = -__logf(u01(w));
      }
      key = f - __logf(q);
    }
    acc_add(a, f, key, i);
  }
  // the per-thread column order is increasing but threads interleave: merge with index tie-break
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    Acc b;
    b.m = __shfl_xor_sync(0xffffffffu, a.m, o);
    b.s = __shfl_xor_sync(0xffffffffu, a.s, o);
    b.key = __shfl_xor_sync(0xffffffffu, a.key, o);
    b.idx = __shfl_xor_sync(0xffffffffu, a.idx, o);
    acc_merge(a, b);
  }
  if (lane == 0) red[warp] = a;
  __syncthreads();
  if (warp == 0) {
    Acc b = lane < kThreads / 32 ? red[lane] : Acc{-INFINITY, 0.f, -INFINITY, 0x7fffffff};
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      Acc c;
      c.m = __shfl_xor_sync(0xffffffffu, b.m, o);
      c.s = __shfl_xor_sync(0xffffffffu, b.s, o);
      c.key = __shfl_xor_sync(0xffffffffu, b.key, o);
      c.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
      acc_merge(b, c);
    }
    if (lane == 0) fin = b;
  }
  __syncthreads();
  const Acc f = fin;
  const int tok = f.idx < vocab ? f.idx : 0;
  const float lse = f.m + __logf(f.s);
  const float xc = scaled<T>(x[tok], t, unit);

  // ---- sweep 2 (L2): rank of the chosen token, optional full probs / logprobs ---------------------
  int above = 0;
  const bool full = out_probs != nullptr || out_logprobs != nullptr;
  float* pr = out_probs != nullptr ? out_probs + (int64_t)row * vocab : nullptr;
  float* lp = out_logprobs != nullptr ? out_logprobs + (int64_t)row * vocab : nullptr;
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    Vec16<T> c;
    *reinterpret_cast<uint4*>(&c) = *reinterpret_cast<const uint4*>(x + (int64_t)v * N);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float g = scaled<T>(c.v[j], t, unit);
      above += g > xc;
      if (full) {
        const float l = g - lse;
        if (lp != nullptr) lp[v * N + j] = l;
        if (pr != nullptr) pr[v * N + j] = __expf(l);
      }
    }
  }
  for (int i = nvec * N + threadIdx.x; i < vocab; i += kThreads) {
    const float g = scaled<T>(x[i], t, unit);
    above += g > xc;
    if (full) {
      const float l = g - lse;
      if (lp != nullptr) lp[i] = l;
      if (pr != nullptr) pr[i] = __expf(l);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
  if (lane == 0) cnt[warp] = above;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < kThreads / 32; ++w) total += cnt[w];
    out_token[row] = tok;
    if (out_logprob != nullptr) out_logprob[row] = xc - lse;
    if (out_rank != nullptr) out_rank[row] = (int64_t)total + 1;
  }
}

}  // namespace sampler
}  // namespace vltr

extern "C" int vltr_sample(const void* logits, int64_t row_stride, int num_rows, int vocab,
                           int dtype, const float* temperatures, const uint8_t* random,
                           const float* noise, uint64_t seed, uint64_t offset, int64_t* out_token,
                           float* out_logprob, int64_t* out_rank, float* out_probs,
                           float* out_logprobs, void* stream) {
  using namespace vltr;
  VLTR_CHECK(num_rows >= 0 && vocab > 0, VLTR_ERR_INVALID, "sample: bad sizes");
  VLTR_CHECK(out_token != nullptr, VLTR_ERR_INVALID, "sample: out_token is required");
  if (num_rows == 0) return VLTR_OK;
  VLTR_DISPATCH_FLOAT(dtype, "sample", {
    VLTR_CHECK(((uintptr_t)logits % sizeof(scalar_t)) == 0, VLTR_ERR_INVALID,
               "sample: misaligned logits");
    launch(sampler::sample_kernel<scalar_t>, dim3(num_rows), dim3(sampler::kThreads), 0,
           (cudaStream_t)stream, (const scalar_t*)logits, row_stride, vocab, temperatures, random,
           noise, seed, offset, out_token, out_logprob, out_rank, out_probs, out_logprobs);
    VLTR_LAUNCH_CHECK("sample");
  });
  return VLTR_OK;
}
