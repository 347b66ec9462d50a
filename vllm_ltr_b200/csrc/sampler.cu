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
// Random sampling, two forms with the same distribution:
//   * caller-supplied noise (`noise` [R, V] fp32 Exp(1) samples): the reference's exponential
//     race, argmax_i p_i / q_i == argmax_i (x_i - log q_i) - what the parity tests run, against
//     the reference fed the same q;
//   * no noise given (production): ONE uniform per row from Philox4x32-10 keyed by (seed, row,
//     offset) and inversion of the row's cumulative distribution - per-thread partial sums from
//     the sweep are scanned across the CTA, the thread whose range holds u * sum is found, and a
//     warp resolves the position inside that thread's elements.  No per-element random numbers
//     or logarithms: the sweep costs one exp per logit, like the greedy one.  (A race over V
//     draws per row measured 4-5x slower than the memory sweep: Philox rounds and two MUFU
//     logarithms per logit.)
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
  const bool race = rnd && nz != nullptr;      // exponential race over the caller's noise
  const bool invert = rnd && nz == nullptr;    // inverse-CDF draw with one in-kernel uniform
  const bool vec_ok = (((uintptr_t)x) % 16 == 0);
  const int nvec = vec_ok ? vocab / N : 0;
  // ---- sweep 1: HBM -> (max, sum, arg-max of the key) ------------------------------------------
  // four 16-byte loads in flight per thread; per vector one rescale of the running sum (when the
  // vector raises the maximum) and N independent exps - no per-element dependency chain
  constexpr int U = 4;
  for (int v0 = threadIdx.x; v0 < nvec; v0 += U * kThreads) {
    Vec16<T> c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kThreads;
      if (v < nvec) *reinterpret_cast<uint4*>(&c[u]) = ld_nc_16(x + (int64_t)v * N);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kThreads;
      if (v < nvec) {
        float f[N];
        float vm = -INFINITY;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          f[j] = scaled<T>(c[u].v[j], t, unit);
          vm = fmaxf(vm, f[j]);
        }
        if (vm > a.m) {
          a.s *= __expf(a.m - vm);          // a.m == -inf: a.s is 0 and stays 0
          a.m = vm;
        }
        if (a.m != -INFINITY) {
          float add = 0.f;
#pragma unroll
          for (int j = 0; j < N; ++j) add += __expf(f[j] - a.m);   // exp(-inf) == 0
          a.s += add;
        }
        if (!invert) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const float key = (race && f[j] != -INFINITY) ? f[j] - __logf(nz[v * N + j]) : f[j];
            if (key > a.key) {     // columns arrive in increasing order: strict > keeps the first
              a.key = key;
              a.idx = v * N + j;
            }
          }
        }
      }
    }
  }
  for (int i = nvec * N + threadIdx.x; i < vocab; i += kThreads) {   // tail / unaligned rows
    const float f = scaled<T>(x[i], t, unit);
    const float key = (race && f != -INFINITY) ? f - __logf(nz[i]) : f;
    acc_add(a, f, key, i);
  }
  const float tm = a.m, ts = a.s;      // this thread's own (max, sum): the inversion needs them
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
  __shared__ float wtot[kThreads / 32];
  __shared__ int sel_t;
  __shared__ float sel_lt;
  __shared__ int sel_tok;
  if (invert) {
    // inclusive scan over the threads of  S_t = sum over thread t's logits of exp(x - max)
    const float st = tm == -INFINITY ? 0.f : ts * __expf(tm - f.m);
    float incl = st;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (threadIdx.x == 0) sel_t = -1;
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    float total = 0.f, before = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      if (w < warp) before += wtot[w];
      total += wtot[w];
    }
    // exclusive bound from the SAME sums as the neighbour's inclusive one: the threads' ranges
    // tile [0, total] exactly, so exactly one of them holds the target
    const float prev = __shfl_up_sync(0xffffffffu, incl, 1);
    const float excl = before + (lane == 0 ? 0.f : prev);
    incl += before;
    const uint4 r4 = philox4x32_10(make_uint4((uint32_t)row, 0x5eedu, (uint32_t)offset,
                                              (uint32_t)(offset >> 32)), pkey);
    const float target = u01(r4.x) * total;           // in (0, total]
    if (st > 0.f && excl < target && target <= incl) {
      sel_t = threadIdx.x;
      sel_lt = target - excl;
    }
    __syncthreads();
    if (warp == 0) {
      const int tstar = sel_t;
      const float lt = sel_lt;
      int chosen = -1, lastidx = -1;
      float run = 0.f;
      if (tstar >= 0) {
        // the selected thread's elements in the order it summed them: vectors tstar + k*kThreads,
        // then the scalar tail; lane k takes the k-th of them, 32 at a time
        const int viters = nvec > tstar ? (nvec - tstar + kThreads - 1) / kThreads : 0;
        for (int k0 = 0; k0 < viters && chosen < 0; k0 += 32) {
          const int k = k0 + lane;
          const bool ok = k < viters;
          const int v = tstar + k * kThreads;
          float e[N];
          float part = 0.f;
          int mylast = -1;
          if (ok) {
            Vec16<T> c;
            *reinterpret_cast<uint4*>(&c) = *reinterpret_cast<const uint4*>(x + (int64_t)v * N);
#pragma unroll
            for (int j = 0; j < N; ++j) {
              const float g = scaled<T>(c.v[j], t, unit);
              e[j] = g == -INFINITY ? 0.f : __expf(g - f.m);
              part += e[j];
              if (e[j] > 0.f) mylast = v * N + j;
            }
          }
          float inc = part;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
          }
          const unsigned nzm = __ballot_sync(0xffffffffu, ok && part > 0.f);
          const unsigned cross = __ballot_sync(0xffffffffu, ok && part > 0.f && run + inc >= lt);
          if (cross != 0u) {
            const int src = __ffs(cross) - 1;
            int cand = -1;
            if (lane == src) {
              float r = run + inc - part;
#pragma unroll
              for (int j = 0; j < N; ++j) {
                if (e[j] > 0.f && cand < 0) {
                  r += e[j];
                  if (r >= lt) cand = v * N + j;
                }
              }
              if (cand < 0) cand = mylast;
            }
            chosen = __shfl_sync(0xffffffffu, cand, src);
          } else if (nzm != 0u) {
            lastidx = __shfl_sync(0xffffffffu, mylast, 31 - __clz(nzm));
          }
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        const int t0 = nvec * N + tstar;
        const int titers = vocab > t0 ? (vocab - t0 + kThreads - 1) / kThreads : 0;
        for (int k0 = 0; k0 < titers && chosen < 0; k0 += 32) {
          const int k = k0 + lane;
          const bool ok = k < titers;
          const int i = t0 + k * kThreads;
          float part = 0.f;
          if (ok) {
            const float g = scaled<T>(x[i], t, unit);
            part = g == -INFINITY ? 0.f : __expf(g - f.m);
          }
          float inc = part;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
          }
          const unsigned nzm = __ballot_sync(0xffffffffu, ok && part > 0.f);
          const unsigned cross = __ballot_sync(0xffffffffu, ok && part > 0.f && run + inc >= lt);
          if (cross != 0u) {
            chosen = t0 + (k0 + __ffs(cross) - 1) * kThreads;
          } else if (nzm != 0u) {
            lastidx = t0 + (k0 + 31 - __clz(nzm)) * kThreads;
          }
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (chosen < 0) chosen = lastidx;      // partial sums rounded below the thread's total
      }
      if (lane == 0) sel_tok = chosen;
    }
    __syncthreads();
  }
  const int tok = invert ? (sel_tok >= 0 && sel_tok < vocab ? sel_tok : (f.idx < vocab ? f.idx : 0))
                         : (f.idx < vocab ? f.idx : 0);
  const float lse = f.m + __logf(f.s);
  const float xc = scaled<T>(x[tok], t, unit);

  // ---- sweep 2 (L2): rank of the chosen token, optional full probs / logprobs ---------------------
  int above = 0;
  const bool full = out_probs != nullptr || out_logprobs != nullptr;
  float* pr = out_probs != nullptr ? out_probs + (int64_t)row * vocab : nullptr;
  float* lp = out_logprobs != nullptr ? out_logprobs + (int64_t)row * vocab : nullptr;
  for (int v0 = threadIdx.x; v0 < nvec; v0 += U * kThreads) {
    Vec16<T> c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kThreads;
      if (v < nvec)
        *reinterpret_cast<uint4*>(&c[u]) = *reinterpret_cast<const uint4*>(x + (int64_t)v * N);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int v = v0 + u * kThreads;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float g = scaled<T>(c[u].v[j], t, unit);
          above += g > xc;
          if (full) {
            const float l = g - lse;
            if (lp != nullptr) lp[v * N + j] = l;
            if (pr != nullptr) pr[v * N + j] = __expf(l);
          }
        }
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
