// ltr_rank.cu — the queue sort of the learning-to-rank scheduling tick.
//
// Replaces the Python `sorted(waiting + running + swapped, key=...)` of the reference's
// Scheduler._get_opt_ordered_requests (vllm/core/scheduler.py:996 with starvation control:
// key = (pri, -aux_model_score); :998 without: key = -aux_model_score) and the sibling
// orderings built the same way (:940-967 tpt/rtpt use (-score, request_id)).
//
// Contract (bit-exact with Python's stable sort): order[] lists the input positions by ascending
// (pri, key), equal keys keep their input order.  Keys are IEEE doubles compared the way Python
// compares floats: -0.0 == +0.0 (tie), everything else by value.  NaN keys are not ordered by
// Python either; here they sort after +inf (positive NaN) / before -inf (negative NaN).
//
// B200 shape: N is the number of live requests (256 .. 10^4 in the reference's benchmarks), far
// too small for a multi-pass radix sort to pay for its launches.  Every request computes its own
// final rank by counting the requests that precede it — rank(i) = #{j : (p_j,k_j) < (p_i,k_i)} +
// #{j < i : (p_j,k_j) == (p_i,k_i)} — with the (pri, key) table staged through shared memory.
// That is N^2 integer compares spread over (N/256) x S CTAs (grid sized to ~2 CTAs per SM):
// one pass, no inter-CTA ordering, stable by construction, 10^8 compares = tens of microseconds
// at N = 10^4.  The last CTA to finish (atomic ticket) turns ranks into the permutation, so the
// whole sort is two launches (key preparation + counting/scatter).
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace vltr {
namespace ltr {

constexpr int kThreads = 256;
constexpr int kTile = 1024;  // (pri, key) entries staged per shared-memory tile

// monotone map double -> uint64 (after canonicalising -0.0 to +0.0)
__device__ __forceinline__ unsigned long long orderable(double v) {
  if (v == 0.0) v = 0.0;  // -0.0 -> +0.0: Python treats them as equal
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// keys[i] = orderable(negate ? -score[i] : score[i]); ranks zeroed
__global__ void __launch_bounds__(kThreads)
    rank_prepare_kernel(const double* __restrict__ score, int negate,
                        unsigned long long* __restrict__ keys, int* __restrict__ ranks, int n,
                        int* __restrict__ ticket) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *ticket = 0;
  if (i >= n) return;
  const double s = score[i];
  keys[i] = orderable(negate ? -s : s);
  ranks[i] = 0;
}

// grid (ceil(n/kThreads), S): CTA (bx, by) counts, for its 256 requests i, the predecessors j in
// the by-th slice of the table.
__global__ void __launch_bounds__(kThreads)
    rank_count_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ pri,
                      int* __restrict__ ranks, int* __restrict__ order, int n, int slice,
                      int* __restrict__ ticket) {
  __shared__ unsigned long long sk[kTile];
  __shared__ int sp[kTile];
  __shared__ int s_last;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  const bool live = i < n;
  const unsigned long long ki = live ? keys[i] : 0ull;
  const int pi = (live && pri != nullptr) ? pri[i] : 0;
  const int j0 = blockIdx.y * slice;
  const int j1 = min(j0 + slice, n);
  int cnt = 0;
  for (int base = j0; base < j1; base += kTile) {
    const int m = min(kTile, j1 - base);
    __syncthreads();
    for (int t = threadIdx.x; t < m; t += kThreads) {
      sk[t] = keys[base + t];
      sp[t] = pri != nullptr ? pri[base + t] : 0;
    }
    __syncthreads();
    if (live) {
      // j < i inside this tile <=> t < i - base
      const int lim = i - base;
#pragma unroll 8
      for (int t = 0; t < m; ++t) {
        const unsigned long long kj = sk[t];
        const int pj = sp[t];
        const bool lt = (pj < pi) || (pj == pi && kj < ki);
        const bool eq = (pj == pi) && (kj == ki);
        cnt += (lt || (eq && t < lim)) ? 1 : 0;
      }
    }
  }
  if (live && cnt != 0) atomicAdd(&ranks[i], cnt);

  // last CTA out scatters the permutation
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = gridDim.x * gridDim.y;
    s_last = (atomicAdd(ticket, 1) == total - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int t = threadIdx.x; t < n; t += kThreads) {
      const int r = __ldcg(&ranks[t]);
      order[r] = t;
    }
  }
}

// ---- per-device staging for the host-pointer entry point -------------------------------------
struct Stage {
  double* h_score = nullptr;  // pinned
  int* h_pri = nullptr;
  int* h_order = nullptr;
  double* d_score = nullptr;
  int* d_pri = nullptr;
  int* d_order = nullptr;
  unsigned long long* d_keys = nullptr;
  int* d_ranks = nullptr;
  int* d_ticket = nullptr;
  int cap = 0;
};
static Stage g_stage[64];
static std::mutex g_mu;

static int ensure(Stage& st, int n) {
  if (n <= st.cap) return VLTR_OK;
  int cap = st.cap ? st.cap : 4096;
  while (cap < n) cap *= 2;
  if (st.cap) {
    VLTR_CUDA(cudaDeviceSynchronize());
    cudaFreeHost(st.h_score); cudaFreeHost(st.h_pri); cudaFreeHost(st.h_order);
    cudaFree(st.d_score); cudaFree(st.d_pri); cudaFree(st.d_order);
    cudaFree(st.d_keys); cudaFree(st.d_ranks);
    st.cap = 0;
  }
  VLTR_CUDA(cudaMallocHost(&st.h_score, sizeof(double) * cap));
  VLTR_CUDA(cudaMallocHost(&st.h_pri, sizeof(int) * cap));
  VLTR_CUDA(cudaMallocHost(&st.h_order, sizeof(int) * cap));
  VLTR_CUDA(cudaMalloc(&st.d_score, sizeof(double) * cap));
  VLTR_CUDA(cudaMalloc(&st.d_pri, sizeof(int) * cap));
  VLTR_CUDA(cudaMalloc(&st.d_order, sizeof(int) * cap));
  VLTR_CUDA(cudaMalloc(&st.d_keys, sizeof(unsigned long long) * cap));
  VLTR_CUDA(cudaMalloc(&st.d_ranks, sizeof(int) * cap));
  if (st.d_ticket == nullptr) VLTR_CUDA(cudaMalloc(&st.d_ticket, sizeof(int)));
  st.cap = cap;
  return VLTR_OK;
}

static int launch(const double* d_score, const int* d_pri, int n, int negate, int* d_order,
                  unsigned long long* d_keys, int* d_ranks, int* d_ticket, cudaStream_t stream) {
  const int bx = (n + kThreads - 1) / kThreads;
  rank_prepare_kernel<<<bx, kThreads, 0, stream>>>(d_score, negate, d_keys, d_ranks, n, d_ticket);
  VLTR_LAUNCH_CHECK("ltr_rank_prepare");
  // slices: aim at ~2 CTAs per SM, at least one tile per slice
  int S = (2 * num_sms() + bx - 1) / bx;
  const int max_s = (n + kTile - 1) / kTile;
  if (S > max_s) S = max_s;
  if (S < 1) S = 1;
  int slice = (n + S - 1) / S;
  slice = ((slice + kTile - 1) / kTile) * kTile;
  S = (n + slice - 1) / slice;
  rank_count_kernel<<<dim3(bx, S), kThreads, 0, stream>>>(d_keys, d_pri, d_ranks, d_order, n,
                                                          slice, d_ticket);
  VLTR_LAUNCH_CHECK("ltr_rank_count");
  return VLTR_OK;
}

}  // namespace ltr
}  // namespace vltr

extern "C" int vltr_ltr_rank(const double* score, const int32_t* pri, int n, int negate,
                             int32_t* order, void* workspace, size_t workspace_bytes,
                             void* stream) {
  using namespace vltr;
  VLTR_CHECK(n >= 0, VLTR_ERR_INVALID, "ltr_rank: negative n");
  if (n == 0) return VLTR_OK;
  const size_t need = vltr_ltr_rank_workspace_bytes(n);
  VLTR_CHECK(workspace != nullptr && workspace_bytes >= need, VLTR_ERR_INVALID,
             "ltr_rank: workspace too small (%zu < %zu)", workspace_bytes, need);
  char* w = (char*)workspace;
  unsigned long long* keys = (unsigned long long*)w;
  int* ranks = (int*)(w + sizeof(unsigned long long) * (size_t)n);
  int* ticket = ranks + n;
  return ltr::launch(score, pri, n, negate, order, keys, ranks, ticket, (cudaStream_t)stream);
}

extern "C" size_t vltr_ltr_rank_workspace_bytes(int n) {
  return sizeof(unsigned long long) * (size_t)n + sizeof(int) * ((size_t)n + 4);
}

extern "C" int vltr_ltr_rank_host(const double* score, const int32_t* pri, int n, int negate,
                                  int32_t* order, void* stream) {
  using namespace vltr;
  VLTR_CHECK(n >= 0, VLTR_ERR_INVALID, "ltr_rank: negative n");
  if (n == 0) return VLTR_OK;
  int dev = 0;
  VLTR_CUDA(cudaGetDevice(&dev));
  VLTR_CHECK(dev >= 0 && dev < 64, VLTR_ERR_INVALID, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(ltr::g_mu);
  ltr::Stage& st = ltr::g_stage[dev];
  int rc = ltr::ensure(st, n);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  memcpy(st.h_score, score, sizeof(double) * n);
  VLTR_CUDA(cudaMemcpyAsync(st.d_score, st.h_score, sizeof(double) * n, cudaMemcpyHostToDevice, s));
  if (pri != nullptr) {
    memcpy(st.h_pri, pri, sizeof(int) * n);
    VLTR_CUDA(cudaMemcpyAsync(st.d_pri, st.h_pri, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  }
  rc = ltr::launch(st.d_score, pri != nullptr ? st.d_pri : nullptr, n, negate, st.d_order,
                   st.d_keys, st.d_ranks, st.d_ticket, s);
  if (rc) return rc;
  VLTR_CUDA(cudaMemcpyAsync(st.h_order, st.d_order, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  VLTR_CUDA(cudaStreamSynchronize(s));
  memcpy(order, st.h_order, sizeof(int) * n);
  return VLTR_OK;
}
