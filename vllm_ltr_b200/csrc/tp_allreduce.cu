// tp_allreduce.cu — the tensor-parallel exchange step of the decode path, fused with what follows it.
//
// In the reference every row-parallel GEMM output ([tokens, hidden]; o_proj and down_proj, two per
// decoder layer) goes through tensor_model_parallel_all_reduce (vllm/model_executor/layers/
// linear.py:577 -> vllm/distributed/communication_op.py:13-42 -> csrc/custom_all_reduce.cuh:181-250
// one-/two-stage kernels over CUDA-IPC pointers, or NCCL) and is then consumed by
// fused_add_rms_norm (vllm/model_executor/models/llama.py:236-238).  Two kernels + a staging copy
// (custom_all_reduce.cu:119-120) per exchange, 160 exchanges per 70B step.
//
// Here the exchange and the norm are ONE kernel over NVLink peer memory:
//   * x (the GEMM output), out and residual live in symmetric buffers (same offset on every rank,
//     every rank holds peer pointers - torch.distributed._symmetric_memory does the plumbing);
//   * rank r owns token rows r, r+world, ...: it sums that row over all ranks reading each peer's
//     copy directly (128-bit loads over NVLink, fp32 accumulation in fixed rank order -> every
//     rank would compute identical bits, like the reference's kernel), rounds to T (the
//     all-reduce result is a T tensor), adds the residual, normalises, and stores the new
//     residual row and the normalised row into EVERY rank's buffers (peer stores);
//   * so each row is reduced and normalised exactly once in the whole group (world x less norm
//     work than the replicated fused_add_rms_norm), and per GPU ~T*H*s*(world-1)/world bytes come
//     in and twice that go out - no intermediate buffer, no second launch;
//   * CTAs of the same index on all ranks rendezvous through flags in the peers' signal pads
//     (system-scope CAS: 0->1 by the sender, 1->0 by the receiver, self-resetting) before reading
//     peer rows and after the peer stores; spins are bounded and trap instead of hanging.
// The plain all-reduce (no norm) uses the same machinery for callers that need it.
#include "common.cuh"

namespace vltr {
namespace tp {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 64;           // barrier channels available in the signal pad
constexpr int kPadBase = 1024;         // our flags start here (u32 index); torch's own barrier uses the front
constexpr int kThreads = 512;

struct Comm {
  const void* x[kMaxWorld];            // every rank's x buffer (index = rank)
  void* out[kMaxWorld];
  void* resid[kMaxWorld];
  uint32_t* pad[kMaxWorld];            // every rank's signal pad
  int rank, world;
};

__device__ __forceinline__ void spin_trap() {
  printf("vltr tp_allreduce: peer barrier timed out\n");
  __trap();
}

// CTA `cta` of every rank meets here.  `release` makes this CTA's prior global / peer stores
// visible system-wide first.
__device__ __forceinline__ void peer_barrier(const Comm& c, int cta, bool release) {
  if (release) __threadfence_system();
  __syncthreads();
  if (threadIdx.x < c.world) {
    const int peer = threadIdx.x;
    uint32_t* remote = c.pad[peer] + kPadBase + cta * kMaxWorld + c.rank;
    uint32_t* local = c.pad[c.rank] + kPadBase + cta * kMaxWorld + peer;
    long long spins = 0;
    while (atomicCAS_system(remote, 0u, 1u) != 0u)
      if (++spins > (1ll << 28)) spin_trap();
    spins = 0;
    while (atomicCAS_system(local, 1u, 0u) != 1u)
      if (++spins > (1ll << 28)) spin_trap();
  }
  __syncthreads();
  __threadfence_system();
}

__device__ __forceinline__ uint4 ld_peer(const void* p) {
  uint4 r;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_peer(void* p, const uint4& v) {
  asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
  return t;
}

// NORM: out = rms_norm(resid += allreduce(x)) ; !NORM: out = allreduce(x)
// VPT: 16-byte vectors per thread (hidden <= VPT * kThreads * 16/sizeof(T))
template <typename T, int VPT, bool NORM>
__global__ void __launch_bounds__(kThreads)
    tp_fused_kernel(const __grid_constant__ Comm c, const T* __restrict__ weight, float eps,
                    int num_tokens, int hidden) {
  constexpr int N = 16 / sizeof(T);
  __shared__ float red[32];
  const int cta = blockIdx.x;
  const int nvec = hidden / N;

  peer_barrier(c, cta, false);      // every rank's x is complete (their previous kernels finished)

  // rows owned by this rank: r, r + world, ...; CTAs of this rank stride over them
  for (int row = c.rank + cta * c.world; row < num_tokens; row += gridDim.x * c.world) {
    const int64_t off = (int64_t)row * hidden;
    float acc[VPT][N];
    Vec16<T> z[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
      for (int j = 0; j < N; ++j) acc[k][j] = 0.f;
    // sum over ranks in rank order (loads of all ranks issued back to back)
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * kThreads;
      if (v < nvec) {
        Vec16<T> part[kMaxWorld];
#pragma unroll
        for (int p = 0; p < kMaxWorld; ++p)
          if (p < c.world)
            *reinterpret_cast<uint4*>(&part[p]) =
                ld_peer(reinterpret_cast<const T*>(c.x[p]) + off + (int64_t)v * N);
#pragma unroll
        for (int p = 0; p < kMaxWorld; ++p)
          if (p < c.world)
#pragma unroll
            for (int j = 0; j < N; ++j) acc[k][j] += ST<T>::to_f(part[p].v[j]);
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * kThreads;
      if (v < nvec) {
        if constexpr (NORM) {
          Vec16<T> r;
          *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(
              reinterpret_cast<const T*>(c.resid[c.rank]) + off + (int64_t)v * N);
#pragma unroll
          for (int j = 0; j < N; ++j) {
            // all-reduce result rounded to T, then the residual add in T (fused_add_rms_norm)
            const float f = rt<T>(acc[k][j]) + ST<T>::to_f(r.v[j]);
            z[k].v[j] = ST<T>::from_f(f);
            const float zf = ST<T>::to_f(z[k].v[j]);
            ss += zf * zf;
          }
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) z[k].v[j] = ST<T>::from_f(acc[k][j]);
        }
      }
    }
    float s = 1.f;
    if constexpr (NORM) {
      ss = block_sum(ss, red);
      s = rsqrtf(ss / hidden + eps);
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * kThreads;
      if (v < nvec) {
        Vec16<T> o;
        if constexpr (NORM) {
          Vec16<T> w;
          *reinterpret_cast<uint4*>(&w) = *reinterpret_cast<const uint4*>(weight + (int64_t)v * N);
#pragma unroll
          for (int j = 0; j < N; ++j)
            o.v[j] = ST<T>::from_f(rt<T>(ST<T>::to_f(z[k].v[j]) * s) * ST<T>::to_f(w.v[j]));
        } else {
          o = z[k];
        }
        // publish the row to every rank
#pragma unroll
        for (int p = 0; p < kMaxWorld; ++p) {
          if (p < c.world) {
            st_peer(reinterpret_cast<T*>(c.out[p]) + off + (int64_t)v * N,
                    *reinterpret_cast<uint4*>(&o));
            if constexpr (NORM)
              st_peer(reinterpret_cast<T*>(c.resid[p]) + off + (int64_t)v * N,
                      *reinterpret_cast<uint4*>(&z[k]));
          }
        }
      }
    }
  }
  peer_barrier(c, cta, true);       // all rows of all ranks have landed everywhere
}

template <typename T, bool NORM>
static int launch(const Comm& c, const void* weight, float eps, int num_tokens, int hidden,
                  cudaStream_t stream) {
  constexpr int N = 16 / sizeof(T);
  const int nvec = hidden / N;
  const int rows_mine = (num_tokens - c.rank + c.world - 1) / c.world;
  // every rank must launch the same grid (CTA b meets CTA b of the peers)
  int grid = (num_tokens + c.world - 1) / c.world;
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (grid < 1) grid = 1;
  (void)rows_mine;
  const int vpt = (nvec + kThreads - 1) / kThreads;
  if (vpt <= 1)
    tp_fused_kernel<T, 1, NORM><<<grid, kThreads, 0, stream>>>(c, (const T*)weight, eps, num_tokens, hidden);
  else if (vpt <= 2)
    tp_fused_kernel<T, 2, NORM><<<grid, kThreads, 0, stream>>>(c, (const T*)weight, eps, num_tokens, hidden);
  else if (vpt <= 4)
    tp_fused_kernel<T, 4, NORM><<<grid, kThreads, 0, stream>>>(c, (const T*)weight, eps, num_tokens, hidden);
  else
    return fail(VLTR_ERR_UNSUPPORTED, "tp_allreduce: hidden size %d too large", hidden);
  VLTR_LAUNCH_CHECK(NORM ? "tp_allreduce_add_rms_norm" : "tp_allreduce");
  return VLTR_OK;
}

static int run(const vltr_tp_comm* comm, const void* weight, float eps, int num_tokens, int hidden,
               int dtype, bool norm, cudaStream_t stream) {
  VLTR_CHECK(comm != nullptr, VLTR_ERR_INVALID, "tp_allreduce: null communicator");
  VLTR_CHECK(comm->world >= 2 && comm->world <= kMaxWorld, VLTR_ERR_UNSUPPORTED,
             "tp_allreduce: world size %d (2..8)", comm->world);
  VLTR_CHECK(comm->rank >= 0 && comm->rank < comm->world, VLTR_ERR_INVALID,
             "tp_allreduce: invalid rank");
  VLTR_CHECK(num_tokens >= 0 && hidden > 0, VLTR_ERR_INVALID, "tp_allreduce: bad sizes");
  VLTR_CHECK(dtype == VLTR_F16 || dtype == VLTR_BF16 || dtype == VLTR_F32, VLTR_ERR_UNSUPPORTED,
             "Unsupported data type: %d", dtype);
  const int n = dtype == VLTR_F32 ? 4 : 8;
  VLTR_CHECK(hidden % n == 0, VLTR_ERR_UNSUPPORTED,
             "tp_allreduce: hidden %d must be a multiple of %d", hidden, n);
  if (num_tokens == 0) return VLTR_OK;
  Comm c{};
  c.rank = comm->rank;
  c.world = comm->world;
  for (int p = 0; p < comm->world; ++p) {
    c.x[p] = comm->x_ptrs[p];
    c.out[p] = comm->out_ptrs[p];
    c.resid[p] = norm ? comm->resid_ptrs[p] : nullptr;
    c.pad[p] = comm->signal_pads[p];
    VLTR_CHECK(c.x[p] != nullptr && c.out[p] != nullptr && c.pad[p] != nullptr &&
                   (!norm || c.resid[p] != nullptr),
               VLTR_ERR_INVALID, "tp_allreduce: null peer pointer for rank %d", p);
    VLTR_CHECK(((uintptr_t)c.x[p] | (uintptr_t)c.out[p] | (uintptr_t)c.resid[p]) % 16 == 0,
               VLTR_ERR_INVALID, "tp_allreduce: peer buffers must be 16-byte aligned");
  }
  VLTR_DISPATCH_FLOAT(dtype, "tp_allreduce", {
    if (norm) return launch<scalar_t, true>(c, weight, eps, num_tokens, hidden, stream);
    return launch<scalar_t, false>(c, nullptr, 0.f, num_tokens, hidden, stream);
  });
  return VLTR_OK;
}

}  // namespace tp
}  // namespace vltr

extern "C" int vltr_tp_allreduce_add_rms_norm(const vltr_tp_comm* comm, const void* weight,
                                              float epsilon, int num_tokens, int hidden_size,
                                              int dtype, void* stream) {
  return vltr::tp::run(comm, weight, epsilon, num_tokens, hidden_size, dtype, true,
                       (cudaStream_t)stream);
}

extern "C" int vltr_tp_allreduce(const vltr_tp_comm* comm, int num_tokens, int hidden_size,
                                 int dtype, void* stream) {
  return vltr::tp::run(comm, nullptr, 0.f, num_tokens, hidden_size, dtype, false,
                       (cudaStream_t)stream);
}
