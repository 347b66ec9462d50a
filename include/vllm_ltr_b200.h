/*
 * vllm_ltr_b200.h — C ABI of libvllm_ltr_b200.so
 *
 * B200 (sm_100a) implementation of the decode-time hot path of hao-ai-lab/vllm-ltr.
 * Every entry point replaces one function of the reference's pybind surface
 * (csrc/pybind.cpp:6-125, declarations csrc/ops.h, csrc/cache.h, csrc/cuda_utils.h) or one
 * piece of the LTR scheduling tick (vllm/core/scheduler.py, vllm/engine/aux_llm_engine.py).
 * The signatures carry plain pointers, sizes, strides and a cudaStream_t (as void*): no
 * torch types.  The Python module vllm_ltr_b200._C re-creates the reference's tensor-level
 * surface (same names / argument order / error behaviour) on top of these calls.
 *
 * Conventions
 *   - all device pointers are owned by the caller; ops write in place exactly like the
 *     reference ops (out / caches / input+residual / query+key);
 *   - every call returns 0 on success, a VLTR_ERR_* code otherwise; vltr_last_error()
 *     returns a thread-local message (what TORCH_CHECK would have put in the RuntimeError);
 *   - kernels are enqueued on `stream` (the caller passes torch's current stream) of the
 *     current device; nothing synchronises unless stated;
 *   - dtype codes: VLTR_F32 / VLTR_F16 / VLTR_BF16 for activations, kv_fp8 = 0 ("auto": cache
 *     has the activation dtype) or 1 ("fp8": unscaled e5m2 bytes, kv_scale ignored on CUDA,
 *     csrc/quantization/fp8_e5m2_kvcache/quant_utils.cuh:164-192).
 */
#ifndef VLLM_LTR_B200_H_
#define VLLM_LTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLTR_OK 0
#define VLTR_ERR_INVALID 1     /* argument rejected (reference: TORCH_CHECK(false, ...)) */
#define VLTR_ERR_CUDA 2        /* CUDA runtime error at launch */
#define VLTR_ERR_UNSUPPORTED 3 /* head size / block size / dtype outside the reference's grid */

#define VLTR_F32 0
#define VLTR_F16 1
#define VLTR_BF16 2
#define VLTR_U8 3

const char* vltr_last_error(void);
/* "sm_100a;<build id>" — lets callers assert the CUDA library (not a fallback) is loaded. */
const char* vltr_build_info(void);
/* number of kernel launches issued through this library since load (process-wide). */
uint64_t vltr_launch_count(void);
/* development knobs, not part of the reference surface.  "pdl" 0/1: launch the hot-path kernels
 * with programmatic dependent launch (default 1, env VLTR_PDL); "attn_warps", "attn_stages",
 * "attn_ctas", "attn_slots", "attn_cap_tokens", "attn_min_tokens": override the paged-attention
 * launch policy (0 = default); "attn_trace": device pointer receiving per-warp timestamps. */
int vltr_tune(const char* key, long long value);

/* ---- attention: csrc/attention/attention_kernels.cu:741-778 (v1), :935-975 (v2) ------------
 * out   [num_seqs, num_heads, head_size]            (contiguous)
 * query [num_seqs, num_heads, head_size]            (row stride q_stride elements)
 * key_cache   [num_blocks, num_kv_heads, head_size/x, block_size, x], x = 16/sizeof(cache elt)
 * value_cache [num_blocks, num_kv_heads, head_size, block_size]
 *             (kv_block_stride / kv_head_stride in elements; a (block, head) tile is contiguous)
 * block_tables [num_seqs, max_num_blocks_per_seq] int32; context_lens [num_seqs] int32
 * alibi_slopes [num_heads] float or NULL.
 * v1 never needs caller scratch; when it decides to split long contexts it uses a library-owned
 * workspace.  v2 uses the caller's exp_sums/max_logits [num_seqs,num_heads,max_num_partitions]
 * float and tmp_out [num_seqs,num_heads,max_num_partitions,head_size] (activation dtype) with
 * the reference's 512-token partitions, and leaves them filled exactly as the reference does. */
int vltr_paged_attention_v1(void* out, const void* query, const void* key_cache,
                            const void* value_cache, int num_seqs, int num_heads, int head_size,
                            int num_kv_heads, float scale, const int32_t* block_tables,
                            const int32_t* context_lens, int block_size, int max_context_len,
                            int max_num_blocks_per_seq, const float* alibi_slopes,
                            int64_t q_stride, int64_t kv_block_stride, int64_t kv_head_stride,
                            int dtype, int kv_fp8, float kv_scale, void* stream);

int vltr_paged_attention_v2(void* out, float* exp_sums, float* max_logits, void* tmp_out,
                            const void* query, const void* key_cache, const void* value_cache,
                            int num_seqs, int num_heads, int head_size, int num_kv_heads,
                            float scale, const int32_t* block_tables,
                            const int32_t* context_lens, int block_size, int max_context_len,
                            int max_num_blocks_per_seq, const float* alibi_slopes,
                            int64_t q_stride, int64_t kv_block_stride, int64_t kv_head_stride,
                            int dtype, int kv_fp8, float kv_scale, void* stream);

/* ---- prefill attention over the paged cache ("prefix prefill") ---------------------------------
 * Replaces the Triton kernel context_attention_fwd (vllm/attention/ops/prefix_prefill.py:630-755;
 * kernels :10-628), which the reference runs for every prefill chunk under chunked prefill
 * (PagedAttention.forward_prefix, vllm/attention/ops/paged_attn.py:163-195).
 *   query/out [num_tokens, num_heads, head_size], key/value [num_tokens, num_kv_heads, head_size]
 *   (row strides in elements); sequence b owns rows start_loc[b] .. + (seq_lens[b] - ctx_lens[b]);
 *   its first ctx_lens[b] tokens are already in the paged cache (layouts as for paged_attention).
 *   out[i] = softmax(scale * q[i] . K(0 .. ctx + i)) . V(0 .. ctx + i), ALiBi optional.
 * The kernel reads ALL keys from the cache.  cache_has_new_tokens = 1: the caller has already
 * written the chunk's key/value into their slots (the engine does, flash_attn.py:192);
 * 0: the op writes them first (key/value must be given).  dtype VLTR_F16 / VLTR_BF16. */
int vltr_context_attention_fwd(void* out, const void* query, const void* key, const void* value,
                               void* key_cache, void* value_cache, const int32_t* block_tables,
                               const int32_t* start_loc, const int32_t* seq_lens,
                               const int32_t* ctx_lens, int num_seqs, int num_tokens,
                               int num_heads, int num_kv_heads, int head_size, int block_size,
                               int max_num_blocks_per_seq, float scale, const float* alibi_slopes,
                               int64_t q_stride, int64_t k_stride, int64_t v_stride,
                               int64_t o_stride, int64_t kv_block_stride, int64_t kv_head_stride,
                               int dtype, int kv_fp8, int cache_has_new_tokens, void* stream);

/* ---- cache ops: csrc/cache_kernels.cu ------------------------------------------------------
 * reshape_and_cache (:235-276): key/value [num_tokens, num_heads, head_size] with row strides
 * key_stride/value_stride (elements); slot_mapping [num_tokens] int64, slot < 0 skipped. */
int vltr_reshape_and_cache(const void* key, const void* value, void* key_cache,
                           void* value_cache, const int64_t* slot_mapping, int num_tokens,
                           int num_heads, int head_size, int block_size, int x,
                           int64_t key_stride, int64_t value_stride, int dtype, int kv_fp8,
                           float kv_scale, void* stream);

/* ---- rotary embedding fused with the KV-cache append (behind the attention seam) ------------------
 * = vltr_rotary_embedding(positions, query, key, ..., is_neox) followed by
 *   vltr_reshape_and_cache(key, value, key_cache, value_cache, slot_mapping, ...)
 * in ONE launch: replaces the pair ops.rotary_embedding (csrc/pos_encoding_kernels.cu:124-172,
 * called from RotaryEmbedding.forward, vllm/model_executor/layers/rotary_embedding.py:136-158, via
 * llama.py:159) + cache_ops.reshape_and_cache (csrc/cache_kernels.cu:235-276, called from
 * flash_attn.py:192).  query [T, Hq, D] and key [T, Hkv, D] are rotated IN PLACE (row strides in
 * elements), value [T, Hkv, D] is read, the caches have the paged layouts above; padding tokens
 * (slot < 0) are rotated but not cached.  Bit-identical to the two separate calls.  Only NeoX
 * rotation over the whole head (rot_dim == head_size) with 16-byte aligned operands is fused:
 * anything else returns VLTR_ERR_UNSUPPORTED and the caller issues the two ops. */
int vltr_rotary_embedding_and_cache(const int64_t* positions, void* query, void* key,
                                    const void* value, void* key_cache, void* value_cache,
                                    const int64_t* slot_mapping, const void* cos_sin_cache,
                                    int64_t num_tokens, int num_heads, int num_kv_heads,
                                    int head_size, int rot_dim, int is_neox, int64_t query_stride,
                                    int64_t key_stride, int64_t value_stride, int block_size, int x,
                                    int dtype, int kv_fp8, void* stream);

/* copy_blocks (:97-152): key_cache_ptrs/value_cache_ptrs are HOST arrays of num_layers device
 * pointers; block_mapping is a HOST array of num_pairs (src,dst) int64 pairs in the reference's
 * std::map iteration order.  block_bytes = bytes of one block of one layer's K (== V).
 * No host<->device copy and no synchronisation (the reference does three blocking H2D). */
int vltr_copy_blocks(void* const* key_cache_ptrs, void* const* value_cache_ptrs, int num_layers,
                     const int64_t* block_mapping, int num_pairs, int64_t block_bytes,
                     void* stream);

/* swap_blocks (:23-62): src/dst base pointers of one layer's K or V cache on device or host.
 * kind: 0 = D2D, 1 = D2H, 2 = H2D.  block_mapping: HOST array of num_pairs (src,dst) pairs.
 * Pinned (device-mapped) host memory is moved by ONE gather/scatter kernel; pageable host
 * memory takes one cudaMemcpyAsync per run of consecutive blocks. */
int vltr_swap_blocks(const void* src, void* dst, int kind, const int64_t* block_mapping,
                     int num_pairs, int64_t block_bytes, void* stream);
/* One swap of ALL layers in one launch (next row).  CacheEngine.swap_in /
 * swap_out (vllm/worker/cache_engine.py:92-100) call PagedAttention.swap_blocks
 * (vllm/attention/ops/paged_attn.py:195-206) per layer, i.e. swap_blocks 2 x layers times, and the
 * reference issues one cudaMemcpyAsync per block in each (csrc/cache_kernels.cu:50-61).
 * src_ptrs / dst_ptrs: HOST arrays of num_bufs buffer pointers (the K and the V cache of every
 * layer, each [num_blocks, block_bytes]); kind as in vltr_swap_blocks; the same block_mapping
 * applies to every buffer.  Pinned host buffers are gathered / scattered by one kernel over UVA;
 * pageable ones fall back to vltr_swap_blocks per buffer. */
int vltr_swap_blocks_layers(const void* const* src_ptrs, void* const* dst_ptrs, int num_bufs,
                            int kind, const int64_t* block_mapping, int num_pairs,
                            int64_t block_bytes, void* stream);

/* convert_fp8 (:298-339): numel elements; exactly one of src/dst dtype is VLTR_U8. */
int vltr_convert_fp8(const void* src, void* dst, int64_t numel, int src_dtype, int dst_dtype,
                     void* stream);

/* ---- layernorm: csrc/layernorm_kernels.cu:277-352 ---------------------------------------- */
int vltr_rms_norm(void* out, const void* input, const void* weight, float epsilon,
                  int num_tokens, int hidden_size, int dtype, void* stream);
int vltr_fused_add_rms_norm(void* input, void* residual, const void* weight, float epsilon,
                            int num_tokens, int hidden_size, int dtype, void* stream);

/* ---- rotary: csrc/pos_encoding_kernels.cu:124-226 ------------------------------------------
 * cos_sin_cache_offsets NULL -> rotary_embedding, else batched_rotary_embedding. */
int vltr_rotary_embedding(const int64_t* positions, void* query, void* key, int64_t num_tokens,
                          int num_heads, int num_kv_heads, int head_size, int rot_dim,
                          int64_t query_stride, int64_t key_stride, const void* cos_sin_cache,
                          const int64_t* cos_sin_cache_offsets, int is_neox, int dtype,
                          void* stream);

/* ---- activations: csrc/activation_kernels.cu:76-161 ----------------------------------------
 * act: 0 silu_and_mul, 1 gelu_and_mul, 2 gelu_tanh_and_mul (input [tokens, 2d] -> out [tokens,d])
 *      3 gelu_new, 4 gelu_fast                        (input [tokens, d]  -> out [tokens,d]) */
int vltr_activation(void* out, const void* input, int64_t num_tokens, int d, int act, int dtype,
                    void* stream);

/* ---- LTR scheduling tick: queue sort ---------------------------------------------------------
 * Replaces `sorted(waiting + running + swapped, key=lambda r: (r.pri, -r.aux_model_score))`
 * (vllm/core/scheduler.py:996; :998 without starvation control: pri == NULL) of
 * Scheduler._get_opt_ordered_requests, bit-exactly: order[] = input positions by ascending
 * (pri, negate ? -score : score), stable (ties keep input order), doubles compared like Python
 * floats (-0.0 == +0.0).
 * vltr_ltr_rank:      score/pri/order are DEVICE arrays of n entries; workspace is device memory
 *                     of vltr_ltr_rank_workspace_bytes(n); asynchronous on `stream`.
 * vltr_ltr_rank_host: score/pri/order are HOST arrays; stages through library-owned pinned
 *                     buffers, runs the kernel and returns after the permutation is on the host. */
size_t vltr_ltr_rank_workspace_bytes(int n);
int vltr_ltr_rank(const double* score, const int32_t* pri, int n, int negate, int32_t* order,
                  void* workspace, size_t workspace_bytes, void* stream);
int vltr_ltr_rank_host(const double* score, const int32_t* pri, int n, int negate, int32_t* order,
                       void* stream);

/* ---- LTR scheduling tick: predictor forward ---------------------------------------------------
 * Replaces the GPU work of AUXLLMEngine.obtain_aux_scores (vllm/engine/aux_llm_engine.py:332-412):
 * OPTForSequenceClassification.forward + compute_logits (vllm/model_executor/models/opt.py:362-397)
 * on a batch of prompts, prefill only, no KV cache.  All pointers are DEVICE pointers owned by the
 * caller; weights are row-major [out, in] (torch nn.Linear / HF checkpoint layout), dtype f16 or
 * bf16; qkv_w/qkv_b are the q,k,v projections stacked along `out` (3*hidden).
 *   token_ids  [num_tokens] int32    prompts concatenated
 *   positions  [num_tokens] int32    0-based position of each token inside its prompt
 *   cu_seqlens [num_prompts+1] int32 prefix sums of prompt lengths
 *   last_rows  [num_prompts] int32   index of each prompt's last token (cu_seqlens[i+1]-1)
 *   logits     [num_prompts, num_labels] float, each value rounded to the model dtype first (the
 *              reference's logits tensor is f16; scores are logits[:,0].tolist(), opt.py:399-409)
 * Asynchronous on `stream`; workspace of vltr_opt_workspace_bytes(handle, tokens, prompts) bytes,
 * 256-byte aligned. */
#define VLTR_OPT_MAX_LAYERS 48
typedef struct vltr_opt_config {
  int hidden, layers, heads, ffn, vocab, max_pos; /* max_pos = max_position_embeddings          */
  int embed_dim;                                  /* word_embed_proj_dim (== hidden for 125M)    */
  int num_labels;
  int do_layer_norm_before;                       /* 1: OPT-125M style pre-LN, 0: OPT-350M post-LN */
  int final_layer_norm;                           /* decoder.final_layer_norm present             */
  int dtype;                                      /* VLTR_F16 / VLTR_BF16                         */
} vltr_opt_config;
typedef struct vltr_opt_weights {
  const void* embed_tokens;    /* [vocab, embed_dim]     */
  const void* embed_positions; /* [max_pos + 2, hidden]  */
  const void* project_in;      /* [hidden, embed_dim] or NULL */
  const void* project_out;     /* [embed_dim, hidden] or NULL */
  const void* final_ln_w;
  const void* final_ln_b;
  const void* score;           /* [num_labels, embed_dim] */
  const void* ln1_w[VLTR_OPT_MAX_LAYERS];  const void* ln1_b[VLTR_OPT_MAX_LAYERS];
  const void* qkv_w[VLTR_OPT_MAX_LAYERS];  const void* qkv_b[VLTR_OPT_MAX_LAYERS];
  const void* out_w[VLTR_OPT_MAX_LAYERS];  const void* out_b[VLTR_OPT_MAX_LAYERS];
  const void* ln2_w[VLTR_OPT_MAX_LAYERS];  const void* ln2_b[VLTR_OPT_MAX_LAYERS];
  const void* fc1_w[VLTR_OPT_MAX_LAYERS];  const void* fc1_b[VLTR_OPT_MAX_LAYERS];
  const void* fc2_w[VLTR_OPT_MAX_LAYERS];  const void* fc2_b[VLTR_OPT_MAX_LAYERS];
} vltr_opt_weights;
int vltr_opt_create(const vltr_opt_config* cfg, const vltr_opt_weights* weights, void** handle);
void vltr_opt_destroy(void* handle);
size_t vltr_opt_workspace_bytes(const void* handle, int max_tokens, int max_prompts);
int vltr_opt_forward(const void* handle, const int32_t* token_ids, const int32_t* positions,
                     const int32_t* cu_seqlens, const int32_t* last_rows, int num_tokens,
                     int num_prompts, int max_len, void* workspace, size_t workspace_bytes,
                     float* logits, void* stream);
/* The dense layer the predictor is built from, exposed for tests:
 * D[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) + resid[M,N]  (tcgen05.mma, fp32 accumulate in TMEM;
 * bias / resid may be NULL; N % 8 == 0, K % 8 == 0; dtype f16 / bf16). */
int vltr_opt_gemm(const void* A, const void* W, void* D, const void* bias, const void* resid, int M,
                  int N, int K, int relu, int dtype, void* stream);

/* ---- lm_head logits (f3, first half) -------------------------------------------------------------
 * Replaces the matmul of LogitsProcessor._get_logits
 * (vllm/model_executor/layers/logits_processor.py:61-64: `torch.matmul(hidden_states,
 * embedding.t())` + optional bias) for the rows _prune_hidden_states selected (:73-78).
 *   logits[num_rows, vocab_rows] = hidden[num_rows, hidden_size] . weight[vocab_rows, hidden_size]^T
 *                                  (+ bias[vocab_rows])
 * all row-major and contiguous, f16 / bf16, fp32 accumulation in TMEM, ONE rounding of the result
 * (the reference rounds the product, then the sum with the bias).  vocab_rows is the rank's
 * shard of the (padded) vocabulary and only needs to be a multiple of 8; hidden_size % 8 == 0.
 * The tensor-parallel gather (:65) and the vocabulary trim (:67-68) stay with the caller. */
int vltr_lm_head_logits(const void* hidden, const void* weight, const void* bias, void* logits,
                        int num_rows, int vocab_rows, int hidden_size, int dtype, void* stream);

/* ---- tensor-parallel exchange (config 5: 70B at TP=8) ------------------------------------------
 * Replaces tensor_model_parallel_all_reduce (vllm/distributed/communication_op.py:13-42; reference
 * kernels csrc/custom_all_reduce.cuh:181-250, entry csrc/custom_all_reduce.cu:101-122) and, fused
 * with it, the fused_add_rms_norm that consumes every all-reduced row-parallel GEMM output
 * (vllm/model_executor/models/llama.py:236-238, csrc/layernorm_kernels.cu:318-352).
 * x / out / resid are SYMMETRIC buffers: one per rank, same shape [num_tokens, hidden], every
 * rank holds all ranks' device pointers (x_ptrs[p] etc. are HOST arrays of `world` peer-mapped
 * device pointers, e.g. from torch.distributed._symmetric_memory); signal_pads[p] is rank p's
 * zero-initialised flag area (>= 8 KiB; the kernels use u32 indices 1024..1040 of it).  Every rank calls with the same sizes and epoch; one
 * kernel per rank, ranks rendezvous through epoch flags in the signal pads.
 *   vltr_tp_allreduce:               out[p] = sum_q x[q]                      on every rank p
 *   vltr_tp_allreduce_add_rms_norm:  resid[p] = (T)(sum_q x[q]) + resid ; out[p] = rms_norm(resid[p]) * weight
 * fp32 accumulation in rank order, results identical on all ranks. */
typedef struct vltr_tp_comm {
  int rank, world;
  uint32_t epoch;                 /* 0: the kernels keep the exchange count in the signal pad
                                     themselves (launch arguments never change: CUDA-graph safe);
                                     or 2 * k + 1 for the k-th exchange on these pads (k = 0, 1,
                                     ...), the same value on every rank, strictly increasing.
                                     One mode per set of signal pads. */
  const void* const* x_ptrs;
  void* const* out_ptrs;
  void* const* resid_ptrs;        /* may be NULL for vltr_tp_allreduce */
  uint32_t* const* signal_pads;
  /* NVLS (optional): multicast mappings of the SAME symmetric buffers (one virtual address that
   * reaches every rank's copy through the NVSwitch, e.g. symmetric-memory handle.multicast_ptr).
   * All three non-NULL: rows are summed by multimem.ld_reduce in the switch and published by
   * multimem.st; any NULL: `world` peer loads / stores per 16 bytes. */
  const void* x_mc;
  void* out_mc;
  void* resid_mc;
  /* 0: resid[p] is updated on every rank (the reference's replicated fused_add_rms_norm).
   * 1: the new residual row is stored only on the rank that owns the row (row % world == rank) -
   *    the only rank that reads it again (the next exchange adds to it); rows a rank does not own
   *    are left untouched in its resid.  Halves the bytes an exchange publishes. */
  int resid_local;
} vltr_tp_comm;
int vltr_tp_allreduce(const vltr_tp_comm* comm, int num_tokens, int hidden_size, int dtype,
                      void* stream);
int vltr_tp_allreduce_add_rms_norm(const vltr_tp_comm* comm, const void* weight, float epsilon,
                                   int num_tokens, int hidden_size, int dtype, void* stream);

/* ---- sampler (next row after the decode path) ------------------------------
 * Replaces the tensor program of Sampler.forward for greedy / temperature-only random sampling
 * (vllm/model_executor/layers/sampler.py:67 temperature, :77-80 softmax + log_softmax in fp32,
 * :395 greedy argmax, :333-357 _multinomial = argmax(probs / Exp(1) noise), :528-544 _get_ranks,
 * :594-600 logprob of the chosen token) with one pass over the logits.
 * logits [num_rows, vocab] (dtype f32/f16/bf16, row stride in elements); temperatures [num_rows]
 * float (already rounded to the logits dtype, as the reference's temperature tensor is) or NULL
 * (= 1.0); random [num_rows] bytes, 0 = greedy, 1 = random, NULL = all greedy; noise [num_rows,
 * vocab] float Exp(1) samples, or NULL to draw them in the kernel (Philox4x32-10 keyed by seed,
 * counter = (column/4, row, offset)).  Outputs: out_token [num_rows] int64 (required),
 * out_logprob [num_rows] float, out_rank [num_rows] int64, and the full fp32 matrices out_probs /
 * out_logprobs [num_rows, vocab] for callers that need top-k logprobs - each may be NULL. */
int vltr_sample(const void* logits, int64_t row_stride, int num_rows, int vocab, int dtype,
                const float* temperatures, const uint8_t* random, const float* noise,
                uint64_t seed, uint64_t offset, int64_t* out_token, float* out_logprob,
                int64_t* out_rank, float* out_probs, float* out_logprobs, void* stream);

/* ---- cuda utils: csrc/cuda_utils_kernels.cu ---------------------------------------------- */
int vltr_get_device_attribute(int attribute, int device_id);
int vltr_get_max_shared_memory_per_block_device_attribute(int device_id);

#ifdef __cplusplus
}
#endif
#endif /* VLLM_LTR_B200_H_ */
