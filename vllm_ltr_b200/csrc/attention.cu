// attention.cu — C-ABI entry points vltr_paged_attention_v1 / _v2 and the launch policy.
//
// Replaces the launchers of csrc/attention/attention_kernels.cu (:628-778 v1, :809-975 v2) and the
// v1/v2 choice made by the caller in vllm/attention/ops/paged_attn.py:101-110.  Both entry points
// run the same single kernel (attention_kernel.cuh); they differ in where partial results live:
//   v1: no caller scratch.  Contexts are cut into partitions sized so that the (sequence, kv head,
//       partition) items fill the machine; partials of multi-partition sequences live in a
//       library-owned fp32 workspace and are merged inside the kernel by the last finisher.
//   v2: the caller's exp_sums / max_logits / tmp_out with the reference's 512-token partitions,
//       left populated the way the reference leaves them (merged in-kernel as well).
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

// ---- per-device state: work-queue counters, partition tickets, v1 partial workspace -------------
// Lifetime rules (a captured CUDA graph bakes these pointers into its launches, and launches on
// different streams may be in flight together):
//   * buffers are grow-only and a buffer that has ever been handed to a launch is NEVER freed:
//     growing allocates a new one and retires the old one (kept alive for graphs that replay it);
//   * nothing is allocated, and the device is never synchronised, while the calling stream is
//     being captured: a launch that would need to grow there fails with a clear error (run one
//     eager call of that shape first - the engine's and DecodeHotPath.capture()'s warm-up does);
//   * every ticket slot owns its own partial workspace, so up to kTicketSlots v1 launches with
//     partitions can overlap on different streams without sharing scratch.
constexpr int kCounterSlots = 64;   // launches that may be in flight on different streams
constexpr int kTicketSlots = 8;
struct DevState {
  int* counters = nullptr;   // kCounterSlots x 2 ints, zero between launches (self-resetting)
  unsigned next_slot = 0;
  int* tickets[kTicketSlots] = {};   // ticket_cap ints each, all zero between launches
  size_t ticket_cap[kTicketSlots] = {};
  char* ws[kTicketSlots] = {};       // fp32 partials of v1 launches with partitions
  size_t ws_bytes[kTicketSlots] = {};
  int* plan_out = nullptr;           // kTicketSlots x 4 ints: partition size chosen by a split launch
  unsigned next_ticket = 0;
};
static DevState g_dev[64];
static std::mutex g_mu;
constexpr size_t kTicketFloor = 131072;  // 256 sequences x 32 kv heads x 16 passes
constexpr size_t kWsFloor = (size_t)256 * 32 * 17 * (128 + 2) * sizeof(float);

static int get_state(DevState** out) {
  int dev = 0;
  VLTR_CUDA(cudaGetDevice(&dev));
  VLTR_CHECK(dev >= 0 && dev < 64, VLTR_ERR_INVALID, "device index %d out of range", dev);
  DevState* st = &g_dev[dev];
  if (st->counters == nullptr) {
    // everything a launch inside the engine's limits (256 sequences, 32 heads of size 128, 17
    // partitions) can need is allocated here, once, for every slot: ~0.6 GB per device, and no
    // later call - eager, captured or replayed - ever has to grow
    int* c = nullptr;
    VLTR_CUDA(cudaMalloc(&c, sizeof(int) * 2 * kCounterSlots));
    VLTR_CUDA(cudaMemset(c, 0, sizeof(int) * 2 * kCounterSlots));
    for (int slot = 0; slot < kTicketSlots; ++slot) {
      VLTR_CUDA(cudaMalloc(&st->tickets[slot], sizeof(int) * kTicketFloor));
      VLTR_CUDA(cudaMemset(st->tickets[slot], 0, sizeof(int) * kTicketFloor));
      st->ticket_cap[slot] = kTicketFloor;
      VLTR_CUDA(cudaMalloc(&st->ws[slot], kWsFloor));
      st->ws_bytes[slot] = kWsFloor;
    }
    VLTR_CUDA(cudaMalloc(&st->plan_out, sizeof(int) * 4 * kTicketSlots));
    VLTR_CUDA(cudaMemset(st->plan_out, 0, sizeof(int) * 4 * kTicketSlots));
    VLTR_CUDA(cudaDeviceSynchronize());
    st->counters = c;
  }
  *out = st;
  return VLTR_OK;
}

static int capturing(cudaStream_t stream, bool* yes) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  VLTR_CUDA(cudaStreamIsCapturing(stream, &cs));
  *yes = cs != cudaStreamCaptureStatusNone;
  return VLTR_OK;
}

static int ensure_ws(DevState* st, int slot, size_t bytes, cudaStream_t stream) {
  if (st->ws_bytes[slot] >= bytes) return VLTR_OK;
  bool cap = false;
  int rc = capturing(stream, &cap);
  if (rc) return rc;
  VLTR_CHECK(!cap, VLTR_ERR_INVALID,
             "paged_attention_v1: the partial workspace must grow (%zu bytes) while the stream is "
             "being captured; run one eager call of this shape before capturing", bytes);
  // beyond the pre-allocated floor (larger batches / more heads than the engine's limits): the
  // old buffer is retired alive, never freed
  // Every slot grows now (launches rotate over the slots: a graph captured after ONE eager call
  // of this shape must not meet a slot that still has to grow).
  const size_t want = bytes + bytes / 2;
  for (int s = 0; s < kTicketSlots; ++s) {
    if (st->ws_bytes[s] >= bytes) continue;
    char* fresh = nullptr;
    VLTR_CUDA(cudaMalloc(&fresh, want));
    st->ws[s] = fresh;
    st->ws_bytes[s] = want;
  }
  return VLTR_OK;
}
static int ensure_tickets(DevState* st, int slot, size_t n, cudaStream_t stream) {
  if (st->ticket_cap[slot] >= n) return VLTR_OK;
  bool cap = false;
  int rc = capturing(stream, &cap);
  if (rc) return rc;
  VLTR_CHECK(!cap, VLTR_ERR_INVALID,
             "paged_attention: the ticket table must grow (%zu entries) while the stream is being "
             "captured; run one eager call of this shape before capturing", n);
  const size_t want = n + n / 2;
  for (int s = 0; s < kTicketSlots; ++s) {     // every slot, for the same reason as ensure_ws
    if (st->ticket_cap[s] >= n) continue;
    int* fresh = nullptr;
    VLTR_CUDA(cudaMalloc(&fresh, sizeof(int) * want));
    // plain (legacy-stream) memset: ordered before any later launch on a blocking stream, and
    // followed by a synchronise for non-blocking ones - growth is a first-calls-only event
    VLTR_CUDA(cudaMemset(fresh, 0, sizeof(int) * want));
    st->tickets[s] = fresh;        // the previous table stays allocated (replayed graphs use it)
    st->ticket_cap[s] = want;
  }
  VLTR_CUDA(cudaDeviceSynchronize());
  return VLTR_OK;
}

template <typename T>
static int launch_merge_t(const Params& p, int head_size, int block_size, cudaStream_t stream) {
  const dim3 grid((p.num_seqs * p.num_heads + 3) / 4), block(128);
  switch (head_size) {
#define VLTR_MERGE(DD) \
    case DD: launch(merge_partials_kernel<T, DD>, grid, block, 0, stream, p, block_size); break;
    VLTR_MERGE(64) VLTR_MERGE(80) VLTR_MERGE(96) VLTR_MERGE(112) VLTR_MERGE(128) VLTR_MERGE(256)
#undef VLTR_MERGE
    default: return fail(VLTR_ERR_UNSUPPORTED, "Unsupported head size: %d", head_size);
  }
  VLTR_LAUNCH_CHECK("paged_attention_merge");
  return VLTR_OK;
}
static int launch_merge(const Params& p, int head_size, int block_size, int dtype,
                        cudaStream_t stream) {
  if (dtype == VLTR_F16) return launch_merge_t<__half>(p, head_size, block_size, stream);
  if (dtype == VLTR_BF16) return launch_merge_t<__nv_bfloat16>(p, head_size, block_size, stream);
  return launch_merge_t<float>(p, head_size, block_size, stream);
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
  VLTR_CHECK(block_size == 8 || block_size == 16 || block_size == 32, VLTR_ERR_UNSUPPORTED,
             "Unsupported block size: %d", block_size);
  if (num_seqs == 0) return VLTR_OK;
  const int cs = kv_fp8 ? 1 : (dtype == VLTR_F32 ? 4 : 2);
  VLTR_CHECK(kv_head_stride == (int64_t)head_size * block_size, VLTR_ERR_INVALID,
             "paged_attention: a (block, head) KV tile must be contiguous (kv_head_stride %lld != "
             "head_size*block_size %d)",
             (long long)kv_head_stride, head_size * block_size);
  VLTR_CHECK(((uintptr_t)key_cache % 16 == 0) && ((uintptr_t)value_cache % 16 == 0) &&
                 (kv_block_stride * cs) % 16 == 0 && (kv_head_stride * cs) % 16 == 0,
             VLTR_ERR_INVALID, "paged_attention: KV cache must be 16-byte aligned");
  VLTR_CHECK((uintptr_t)query % 16 == 0 && (q_stride * (dtype == VLTR_F32 ? 4 : 2)) % 16 == 0,
             VLTR_ERR_INVALID, "paged_attention: query rows must be 16-byte aligned");

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
  p.use_plan = num_seqs <= kMaxPlanSeqs;

  // query heads per work item: G == 1 (or fp32) runs the FMA path one head group at a time, GQA
  // groups ride the 16 rows of the tensor-core tile
  const int G = p.group;
  const bool mma = dtype != VLTR_F32 && G >= 2;
  int gt;
  if (mma) gt = G <= 8 ? 8 : 16;
  else gt = G <= 1 ? 1 : (G <= 2 ? 2 : (G <= 4 ? 4 : 8));
  // (the FMA instantiations cap gt by register budget; launch_gt() may lower it)
  p.passes = (G + gt - 1) / gt;

  const int max_blocks = (max_context_len + block_size - 1) / block_size;
  if (v2) {
    // reference partitioning: PARTITION_SIZE = 512 (attention_kernels.cu:815,858)
    p.part_blocks = 512 / block_size;
    p.n_part = (max_context_len + 511) / 512;
    if (p.n_part < 1) p.n_part = 1;
    p.always_partial = 1;
    p.tmp_is_f32 = dtype == VLTR_F32;
    p.exp_sums = exp_sums;
    p.max_logits = max_logits;
    p.tmp_out = tmp_out;
  } else {
    // partition size: enough (sequence, kv head, pass, partition) items to give every resident
    // warp one, at most 512 tokens (one lane-held window of block ids per item, and a short
    // tail), at least 64 tokens, and no more than 16 partitions per sequence (the merge of a
    // sequence is done by one warp)
    const long long slots =
        (long long)num_sms() * (g_tune.attn_slots > 0 ? g_tune.attn_slots : 12);
    const long long floor_tokens = g_tune.attn_min_tokens > 0 ? g_tune.attn_min_tokens : 64;
    const long long cap_tokens = g_tune.attn_cap_tokens > 0 ? g_tune.attn_cap_tokens : 512;
    const long long groups = (long long)num_seqs * num_kv_heads * p.passes;
    long long parts_wanted = (slots + groups - 1) / groups;
    if (parts_wanted < 1) parts_wanted = 1;
    long long pt = (max_context_len + parts_wanted - 1) / parts_wanted;
    const long long pt_min = (max_context_len + 15) / 16;
    if (pt < pt_min) pt = pt_min;
    if (pt < floor_tokens) pt = floor_tokens;
    int cap = 32 * block_size;
    if (cap > cap_tokens) cap = (int)cap_tokens;
    if (pt > cap && pt_min <= cap) pt = cap;
    if (pt > 64 * block_size) pt = 64 * block_size;
    int pb = (int)((pt + block_size - 1) / block_size);
    if (pb < 1) pb = 1;
    if (pb > 64) pb = 64;
    p.part_blocks = pb;
    p.n_part = max_blocks > 0 ? (max_blocks + pb - 1) / pb : 1;
    // (The kernel can also size partitions itself from the actual contexts - Params::adaptive.
    // Measured on B200 it loses to this host policy: finer partitions cost more in per-item
    // set-up and merging than they win in balance, see profiles/r01_op_sweep_adaptive.txt.)
    p.adaptive = 0;
    p.part_blocks_cap = pb;
    p.always_partial = 0;
    p.tmp_is_f32 = 1;
    // Latency-bound regime: not every resident warp would get a (sequence, kv head, pass) group
    // of its own - tensor-parallel shards (70B at TP8: one kv head per GPU), small batches.  The
    // kernel then sizes the partitions itself from the ACTUAL contexts (~1.5 items per resident
    // warp, at least 64 tokens each) and the merge becomes a second, fully parallel launch
    // (merge_partials_kernel) instead of the last finisher's chore.  attn_split: 0 = this policy,
    // 1 = always when possible, -1 = never.
    const bool can_split = p.use_plan && max_blocks > 1;
    const bool want_split = g_tune.attn_split > 0 || (g_tune.attn_split == 0 && groups <= slots);
    if (can_split && want_split) {
      // at least 128 tokens per partition unless tuned (64 measured ~15% slower on the TP8 shards:
      // twice the partials for the merge launch, profiles/r02_attn_split.txt)
      int pb_min = (int)(g_tune.attn_min_tokens > 0 ? g_tune.attn_min_tokens : 128) / block_size;
      if (pb_min < 1) pb_min = 1;
      if (pb_min < (max_blocks + 127) / 128) pb_min = (max_blocks + 127) / 128;   // nparts <= 128
      if (pb_min < pb) {
        p.adaptive = 1;
        p.part_blocks = pb_min;
        p.part_blocks_cap = pb;
        p.n_part = (max_blocks + pb_min - 1) / pb_min;
        p.defer_merge = 1;
      }
    }
  }
  int slot = 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    slot = (int)(st->next_ticket++ % kTicketSlots);
    if (!v2 && p.n_part > 1) {
      const size_t n = (size_t)num_seqs * num_heads * p.n_part;
      const size_t bytes = n * sizeof(float) * 2 + n * head_size * sizeof(float);
      int rc = ensure_ws(st, slot, bytes, stream);
      if (rc) return rc;
      p.exp_sums = reinterpret_cast<float*>(st->ws[slot]);
      p.max_logits = p.exp_sums + n;
      p.tmp_out = p.max_logits + n;
    }
    const size_t ngroups = (size_t)num_seqs * num_kv_heads * 16;   // passes <= 16 for any sane G
    int rc = ensure_tickets(st, slot, ngroups, stream);
    if (rc) return rc;
    p.tickets = st->tickets[slot];
    p.counters = st->counters + 2 * (st->next_slot++ % kCounterSlots);
    p.plan_out = p.defer_merge ? st->plan_out + 4 * slot : nullptr;
  }
  VLTR_CHECK((long long)num_seqs * num_kv_heads * p.passes <= (long long)st->ticket_cap[slot],
             VLTR_ERR_INVALID, "paged_attention: too many head groups");

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
  if (rc != VLTR_OK || !p.defer_merge) return rc;
  return launch_merge(p, head_size, block_size, dtype, stream);
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
