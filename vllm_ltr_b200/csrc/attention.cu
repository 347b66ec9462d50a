// attention.cu — C-ABI entry points vltr_paged_attention_v1 / _v2 and the launch policy.
//
// Replaces the launchers of csrc/attention/attention_kernels.cu (:628-778 v1, :809-975 v2) and the
// v1/v2 choice made by the caller in vllm/attention/ops/paged_attn.py:101-110.  Both entry points
// run the same kernel (attention_kernel.cuh); they differ in where partial results live:
//   v1: no caller scratch.  Contexts are split only when that buys parallelism or trims the tail
//       (few work items, or a long maximum context); partials then live in a library-owned fp32
//       workspace and are merged by a second small kernel.
//   v2: the caller's exp_sums / max_logits / tmp_out with the reference's 512-token partitions,
//       left populated the way the reference leaves them.
#include <mutex>

#include "attention_kernel.cuh"

namespace vltr {
namespace attn {

int launch_mode_f16(Params&, int, int, cudaStream_t);
int launch_mode_bf16(Params&, int, int, cudaStream_t);
int launch_mode_f32(Params&, int, int, cudaStream_t);
int launch_mode_f16_fp8(Params&, int, int, cudaStream_t);
int launch_mode_bf16_fp8(Params&, int, int, cudaStream_t);
int launch_mode_f32_fp8(Params&, int, int, cudaStream_t);
int reduce_f16(void*, const Params&, int, int, cudaStream_t);
int reduce_bf16(void*, const Params&, int, int, cudaStream_t);
int reduce_f32(void*, const Params&, int, int, cudaStream_t);

// ---- per-device state: work-queue counters and the v1 split workspace ---------------------------
constexpr int kCounterSlots = 64;
struct DevState {
  int* counters = nullptr;   // kCounterSlots x 2 ints, zero between launches (self-resetting)
  unsigned next_slot = 0;
  char* ws = nullptr;        // grow-only
  size_t ws_bytes = 0;
};
static DevState g_dev[64];
static std::mutex g_mu;

static int get_state(DevState** out) {
  int dev = 0;
  VLTR_CUDA(cudaGetDevice(&dev));
  VLTR_CHECK(dev >= 0 && dev < 64, VLTR_ERR_INVALID, "device index %d out of range", dev);
  DevState* st = &g_dev[dev];
  if (st->counters == nullptr) {
    VLTR_CUDA(cudaMalloc(&st->counters, sizeof(int) * 2 * kCounterSlots));
    VLTR_CUDA(cudaMemset(st->counters, 0, sizeof(int) * 2 * kCounterSlots));
  }
  *out = st;
  return VLTR_OK;
}

static int ensure_ws(DevState* st, size_t bytes) {
  if (st->ws_bytes >= bytes) return VLTR_OK;
  // grow-only; growing frees the old buffer after the device drains (rare: first calls only)
  if (st->ws != nullptr) {
    VLTR_CUDA(cudaDeviceSynchronize());
    VLTR_CUDA(cudaFree(st->ws));
    st->ws = nullptr;
    st->ws_bytes = 0;
  }
  size_t cap = bytes + bytes / 2;
  VLTR_CUDA(cudaMalloc(&st->ws, cap));
  st->ws_bytes = cap;
  return VLTR_OK;
}

static int run(bool v2, void* out, float* exp_sums, float* max_logits, void* tmp_out,
               const void* query, const void* key_cache, const void* value_cache, int num_seqs,
               int num_heads, int head_size, int num_kv_heads, float scale,
               const int32_t* block_tables, const int32_t* context_lens, int block_size,
               int max_context_len, int max_num_blocks_per_seq, const float* alibi_slopes,
               int64_t q_stride, int64_t kv_block_stride, int64_t kv_head_stride, int dtype,
               int kv_fp8, cudaStream_t stream) {
  VLTR_CHECK(num_seqs >= 0 && num_heads > 0 && num_kv_heads > 0, VLTR_ERR_INVALID,
             "paged_attention: bad sizes");
  VLTR_CHECK(num_heads % num_kv_heads == 0, VLTR_ERR_INVALID,
             "paged_attention: num_heads %d not divisible by num_kv_heads %d", num_heads,
             num_kv_heads);
  VLTR_CHECK(dtype == VLTR_F32 || dtype == VLTR_F16 || dtype == VLTR_BF16, VLTR_ERR_UNSUPPORTED,
             "Unsupported data type: %d", dtype);
  if (num_seqs == 0) return VLTR_OK;
  const int cs = kv_fp8 ? 1 : (dtype == VLTR_F32 ? 4 : 2);
  const int as = dtype == VLTR_F32 ? 4 : 2;
  VLTR_CHECK(kv_head_stride == (int64_t)head_size * block_size, VLTR_ERR_INVALID,
             "paged_attention: a (block, head) KV tile must be contiguous (kv_head_stride %lld != "
             "head_size*block_size %d)",
             (long long)kv_head_stride, head_size * block_size);
  VLTR_CHECK(((uintptr_t)key_cache % 16 == 0) && ((uintptr_t)value_cache % 16 == 0) &&
                 (kv_block_stride * cs) % 16 == 0 && (kv_head_stride * cs) % 16 == 0,
             VLTR_ERR_INVALID, "paged_attention: KV cache must be 16-byte aligned");
  (void)as;

  DevState* st = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = get_state(&st);
    if (rc) return rc;
  }

  Params p{};
  p.out = out;
  p.q = query;
  p.k_cache = key_cache;
  p.v_cache = value_cache;
  p.block_tables = block_tables;
  p.context_lens = context_lens;
  p.alibi_slopes = alibi_slopes;
  p.num_seqs = num_seqs;
  p.num_heads = num_heads;
  p.num_kv_heads = num_kv_heads;
  p.group = num_heads / num_kv_heads;
  p.max_blocks_per_seq = max_num_blocks_per_seq;
  p.scale = scale;
  p.q_stride = q_stride;
  p.kv_block_stride = kv_block_stride * cs;
  p.kv_head_stride = kv_head_stride * cs;

  const int max_blocks = (max_context_len + block_size - 1) / block_size;
  int part_tokens = 0;
  if (v2) {
    // reference partitioning: PARTITION_SIZE = 512 (attention_kernels.cu:815,858)
    VLTR_CHECK(512 % block_size == 0, VLTR_ERR_UNSUPPORTED, "Unsupported block size: %d",
               block_size);
    part_tokens = 512;
    p.n_part = (max_context_len + 511) / 512;
    if (p.n_part < 1) p.n_part = 1;
    p.part_blocks = 512 / block_size;
    p.split = 1;
    p.tmp_is_f32 = dtype == VLTR_F32;
    p.exp_sums = exp_sums;
    p.max_logits = max_logits;
    p.tmp_out = tmp_out;
  } else {
    // split when the (seq, kv_head) items cannot fill the machine, or when one long context
    // would become the tail of the launch
    const int slots = num_sms() * 8;
    const long long items = (long long)num_seqs * num_kv_heads;
    int want = 1;
    if (items < 2LL * slots) want = (int)((2LL * slots + items - 1) / items);
    if (max_context_len > 2048) {
      const int by_len = (max_context_len + 1023) / 1024;
      if (by_len > want) want = by_len;
    }
    int pb = (max_blocks + want - 1) / want;                 // blocks per partition
    const int min_pb = (256 + block_size - 1) / block_size;  // at least 256 tokens
    if (pb < min_pb) pb = min_pb;
    if (pb < 1) pb = 1;
    const int n_part = max_blocks > 0 ? (max_blocks + pb - 1) / pb : 1;
    if (n_part <= 1) {
      p.n_part = 1;
      p.part_blocks = max_blocks > 0 ? max_blocks : 1;
      p.split = 0;
    } else {
      p.n_part = n_part;
      p.part_blocks = pb;
      part_tokens = pb * block_size;
      p.split = 1;
      p.tmp_is_f32 = 1;
      const size_t n = (size_t)num_seqs * num_heads * n_part;
      const size_t bytes = n * sizeof(float) * 2 + n * head_size * sizeof(float);
      {
        std::lock_guard<std::mutex> lk(g_mu);
        int rc = ensure_ws(st, bytes);
        if (rc) return rc;
      }
      p.exp_sums = reinterpret_cast<float*>(st->ws);
      p.max_logits = p.exp_sums + n;
      p.tmp_out = p.max_logits + n;
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    p.counters = st->counters + 2 * (st->next_slot++ % kCounterSlots);
  }

  int rc;
  if (!kv_fp8) {
    rc = dtype == VLTR_F16    ? launch_mode_f16(p, head_size, block_size, stream)
         : dtype == VLTR_BF16 ? launch_mode_bf16(p, head_size, block_size, stream)
                              : launch_mode_f32(p, head_size, block_size, stream);
  } else {
    rc = dtype == VLTR_F16    ? launch_mode_f16_fp8(p, head_size, block_size, stream)
         : dtype == VLTR_BF16 ? launch_mode_bf16_fp8(p, head_size, block_size, stream)
                              : launch_mode_f32_fp8(p, head_size, block_size, stream);
  }
  if (rc) return rc;
  if (p.split) {
    rc = dtype == VLTR_F16    ? reduce_f16(out, p, head_size, part_tokens, stream)
         : dtype == VLTR_BF16 ? reduce_bf16(out, p, head_size, part_tokens, stream)
                              : reduce_f32(out, p, head_size, part_tokens, stream);
  }
  return rc;
}

}  // namespace attn
}  // namespace vltr

extern "C" int vltr_paged_attention_v1(void* out, const void* query, const void* key_cache,
                                       const void* value_cache, int num_seqs, int num_heads,
                                       int head_size, int num_kv_heads, float scale,
                                       const int32_t* block_tables, const int32_t* context_lens,
                                       int block_size, int max_context_len,
                                       int max_num_blocks_per_seq, const float* alibi_slopes,
                                       int64_t q_stride, int64_t kv_block_stride,
                                       int64_t kv_head_stride, int dtype, int kv_fp8,
                                       float kv_scale, void* stream) {
  (void)kv_scale;
  return vltr::attn::run(false, out, nullptr, nullptr, nullptr, query, key_cache, value_cache,
                         num_seqs, num_heads, head_size, num_kv_heads, scale, block_tables,
                         context_lens, block_size, max_context_len, max_num_blocks_per_seq,
                         alibi_slopes, q_stride, kv_block_stride, kv_head_stride, dtype, kv_fp8,
                         (cudaStream_t)stream);
}

extern "C" int vltr_paged_attention_v2(void* out, float* exp_sums, float* max_logits,
                                       void* tmp_out, const void* query, const void* key_cache,
                                       const void* value_cache, int num_seqs, int num_heads,
                                       int head_size, int num_kv_heads, float scale,
                                       const int32_t* block_tables, const int32_t* context_lens,
                                       int block_size, int max_context_len,
                                       int max_num_blocks_per_seq, const float* alibi_slopes,
                                       int64_t q_stride, int64_t kv_block_stride,
                                       int64_t kv_head_stride, int dtype, int kv_fp8,
                                       float kv_scale, void* stream) {
  (void)kv_scale;
  return vltr::attn::run(true, out, exp_sums, max_logits, tmp_out, query, key_cache, value_cache,
                         num_seqs, num_heads, head_size, num_kv_heads, scale, block_tables,
                         context_lens, block_size, max_context_len, max_num_blocks_per_seq,
                         alibi_slopes, q_stride, kv_block_stride, kv_head_stride, dtype, kv_fp8,
                         (cudaStream_t)stream);
}
