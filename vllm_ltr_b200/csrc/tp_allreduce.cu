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
//   * ranks rendezvous before reading peer rows and after the peer stores through monotonically
//     increasing epoch flags in the peers' signal pads: one-way release stores over NVLink and
//     local acquire polls (no round trips, nothing to reset); one CTA per rank talks to the
//     peers, the rest follow a device-scope flag; spins are bounded and trap instead of hanging.
// The plain all-reduce (no norm) uses the same machinery for callers that need it.
//
// NVLS form (comm->x_mc / out_mc / resid_mc non-null: multicast mappings of the same symmetric
// buffers, NVSwitch gen 3+): the owner of a row issues ONE multimem.ld_reduce per 16 bytes - the
// switch fetches the row from all ranks and adds it up (fp32 accumulation) - instead of `world`
// peer loads, and publishes the normalised row and the new residual with ONE multimem.st each
// instead of `world` peer stores.  Same rendezvous, same ownership, same rounding points; the
// summation order inside the switch is fixed by the hardware, and since only the owner reduces a
// row every rank still receives identical bits.
//
// Launch: programmatic dependent launch like the other hot-path kernels - the grid is resident
// and has read its epoch while the producer of x drains; the first cross-rank announcement comes
// after griddepcontrol.wait (it says "my x is complete").
#include "common.cuh"

namespace vltr {
namespace tp {

constexpr int kMaxWorld = 8;
constexpr int kPadBase = 1024;         // our flags start here (u32 index); torch's own barrier uses the front
constexpr int kThreads = 256;
// signal-pad layout (u32 index from kPadBase): [0..7] arrival epoch of rank p (written by p),
// [9] local count of finished CTAs, [10] this rank's exchange counter (device-epoch mode)

struct Comm {
  const void* x[kMaxWorld];            // every rank's x buffer (index = rank)
  void* out[kMaxWorld];
  void* resid[kMaxWorld];
  uint32_t* pad[kMaxWorld];            // every rank's signal pad
  const void* x_mc;                    // multicast mappings (NVLS) or null
  void* out_mc;
  void* resid_mc;
  long long* trace;                    // development: 8 globaltimer stamps per call (CTA 0) or null
  int resid_local;                     // 1: the new residual row is stored on its owner only
  int rank, world;
  uint32_t epoch;                      // 2 * call_index + 1 (start barrier), + 1 (end barrier);
                                       // 0: read it from this rank's counter in the signal pad
};

__device__ __forceinline__ void spin_trap() {
  printf("vltr tp_allreduce: peer barrier timed out\n");
  __trap();
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Ranks meet through epoch flags: a rank ANNOUNCES by writing `epoch` into its slot of every
// peer's pad (one-way NVLink stores, no round trips; release: everything this rank wrote before
// is visible first), and WAITS by polling its own pad until every slot has reached `epoch`.
// Epochs grow monotonically, so flags never need resetting.
__device__ __forceinline__ void announce(const Comm& c, uint32_t epoch) {   // one CTA per rank
  if (threadIdx.x < c.world) st_release_sys(c.pad[threadIdx.x] + kPadBase + c.rank, epoch);
}
__device__ __forceinline__ void wait_peers(const Comm& c, uint32_t epoch) {  // any CTA
  if (threadIdx.x < c.world) {
    const uint32_t* mine = c.pad[c.rank] + kPadBase + threadIdx.x;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0)
      if (++spins > (1ll << 30)) spin_trap();
  }
  __syncthreads();
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

// NVLS: one load that the switch turns into a reduction over every rank's copy
template <typename T>
__device__ __forceinline__ uint4 ld_reduce_mc(const void* p);
template <>
__device__ __forceinline__ uint4 ld_reduce_mc<__nv_bfloat16>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 ld_reduce_mc<__half>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 ld_reduce_mc<float>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_mc(void* p, const uint4& v) {   // one store, every rank's copy
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
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
template <typename T, int VPT, bool NORM, bool MC>
__global__ void __launch_bounds__(kThreads)
    tp_fused_kernel(const __grid_constant__ Comm c, const T* __restrict__ weight, float eps,
                    int num_tokens, int hidden) {
  constexpr int N = 16 / sizeof(T);
  __shared__ float red[32];
  const int cta = blockIdx.x;
  long long* tr = (c.trace != nullptr && cta == 0 && threadIdx.x == 0) ? c.trace : nullptr;
  if (tr) tr[0] = gtimer();
  const int nvec = hidden / N;
  uint32_t* fin = c.pad[c.rank] + kPadBase + 9;
  // Device-epoch mode (CUDA-graph safe: nothing in the launch arguments changes from call to
  // call): the epoch is 1 + this rank's exchange counter, which the last CTA of every exchange
  // advances by 2.  Every rank issues the same exchanges in the same order, so the counters
  // agree.  All CTAs read it before any can finish (the advance waits for all of them).
  uint32_t* ctr = c.pad[c.rank] + kPadBase + 10;
  // (the counter is advanced by the previous exchange of THIS rank's stream: wait for it - and
  // for the producer of x - before reading; under programmatic dependent launch everything above
  // overlapped their tails)
  pdl_wait();
  pdl_trigger();
  const uint32_t epoch = c.epoch != 0 ? c.epoch : *reinterpret_cast<volatile uint32_t*>(ctr) + 1u;
  if (tr) tr[1] = gtimer();

  // start: every rank's x is complete (its producing kernels finished before this one started).
  // CTA 0 announces this rank; every CTA polls the local pad for all peers' announcements.
  if (cta == 0) announce(c, epoch);
  wait_peers(c, epoch);
  if (tr) tr[2] = gtimer();

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
        if constexpr (MC) {
          // the switch adds the row up (fp32 accumulate) and returns it rounded to T
          Vec16<T> sum;
          *reinterpret_cast<uint4*>(&sum) =
              ld_reduce_mc<T>(reinterpret_cast<const T*>(c.x_mc) + off + (int64_t)v * N);
#pragma unroll
          for (int j = 0; j < N; ++j) acc[k][j] = ST<T>::to_f(sum.v[j]);
        } else {
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
        // The residual of a row is only ever read again by the row's owner (the next exchange
        // adds to it): with resid_local it stays here, and half of the published bytes - and of
        // what the system fence below has to wait for - are gone.
        if constexpr (NORM) {
          if (c.resid_local)
            *reinterpret_cast<uint4*>(reinterpret_cast<T*>(c.resid[c.rank]) + off + (int64_t)v * N) =
                *reinterpret_cast<uint4*>(&z[k]);
        }
        // publish the row to every rank
        if constexpr (MC) {
          st_mc(reinterpret_cast<T*>(c.out_mc) + off + (int64_t)v * N, *reinterpret_cast<uint4*>(&o));
          if constexpr (NORM)
            if (!c.resid_local)
              st_mc(reinterpret_cast<T*>(c.resid_mc) + off + (int64_t)v * N,
                    *reinterpret_cast<uint4*>(&z[k]));
        } else {
#pragma unroll
          for (int p = 0; p < kMaxWorld; ++p) {
            if (p < c.world) {
              st_peer(reinterpret_cast<T*>(c.out[p]) + off + (int64_t)v * N,
                      *reinterpret_cast<uint4*>(&o));
              if constexpr (NORM)
                if (!c.resid_local)
                  st_peer(reinterpret_cast<T*>(c.resid[p]) + off + (int64_t)v * N,
                          *reinterpret_cast<uint4*>(&z[k]));
            }
          }
        }
      }
    }
  }
  // end: all rows of all ranks have landed everywhere.  The last CTA of this rank to finish its
  // peer stores announces (one system-scope release by one thread covers the CTA's stores: the
  // CTA barrier orders them before it) and waits for the peers.
  if (tr) tr[3] = gtimer();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (tr) tr[4] = gtimer();
    const uint32_t done = atomicAdd(fin, 1u);
    s_last = done == gridDim.x - 1;
    if (s_last) *fin = 0u;
  }
  __syncthreads();
  if (s_last) {
    announce(c, epoch + 1);
    wait_peers(c, epoch + 1);
    if (c.epoch == 0 && threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(ctr) = epoch + 1u;
  }
  if (tr) {
    tr[5] = gtimer();
    tr[6] = s_last;
  }
}

template <typename T, bool NORM>
static int launch(const Comm& c, const void* weight, float eps, int num_tokens, int hidden,
                  cudaStream_t stream) {
  constexpr int N = 16 / sizeof(T);
  const int nvec = hidden / N;
  const int rows_mine = (num_tokens - c.rank + c.world - 1) / c.world;
  // one CTA per owned row, up to two per SM (they loop beyond that)
  int grid = (num_tokens + c.world - 1) / c.world;
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  if (grid < 1) grid = 1;
  (void)rows_mine;
  const int vpt = (nvec + kThreads - 1) / kThreads;
  const bool mc = c.x_mc != nullptr && g_tune.tp_multimem >= 0;
#define VLTR_TP_LAUNCH(V)                                                                         \
  do {                                                                                            \
    if (mc)                                                                                       \
      vltr::launch(tp_fused_kernel<T, V, NORM, true>, dim3(grid), dim3(kThreads), 0, stream, c,   \
                   (const T*)weight, eps, num_tokens, hidden);                                    \
    else                                                                                          \
      vltr::launch(tp_fused_kernel<T, V, NORM, false>, dim3(grid), dim3(kThreads), 0, stream, c,  \
                   (const T*)weight, eps, num_tokens, hidden);                                    \
  } while (0)
  if (vpt <= 1) VLTR_TP_LAUNCH(1);
  else if (vpt <= 2) VLTR_TP_LAUNCH(2);
  else if (vpt <= 4) VLTR_TP_LAUNCH(4);
  else if (vpt <= 8) VLTR_TP_LAUNCH(8);
  else
    return fail(VLTR_ERR_UNSUPPORTED, "tp_allreduce: hidden size %d too large", hidden);
#undef VLTR_TP_LAUNCH
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
  VLTR_CHECK(comm->epoch == 0 || comm->epoch % 2 == 1, VLTR_ERR_INVALID,
             "tp_allreduce: epoch must be 0 (device-side counter) or odd (2 * call_index + 1), "
             "got %u", comm->epoch);
  Comm c{};
  c.rank = comm->rank;
  c.world = comm->world;
  c.epoch = comm->epoch;
  c.x_mc = comm->x_mc;
  c.out_mc = comm->out_mc;
  c.resid_mc = norm ? comm->resid_mc : nullptr;
  if (c.x_mc == nullptr || c.out_mc == nullptr || (norm && c.resid_mc == nullptr))
    c.x_mc = nullptr;                 // all three multicast mappings or none
  VLTR_CHECK(((uintptr_t)c.x_mc | (uintptr_t)c.out_mc | (uintptr_t)c.resid_mc) % 16 == 0,
             VLTR_ERR_INVALID, "tp_allreduce: multicast mappings must be 16-byte aligned");
  c.trace = reinterpret_cast<long long*>(g_tune.tp_trace);
  c.resid_local = comm->resid_local != 0;
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
