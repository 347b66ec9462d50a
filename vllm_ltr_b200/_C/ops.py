"""``vllm._C.ops`` — attention, activation, layernorm and rotary ops (csrc/pybind.cpp:8-60,
declarations csrc/ops.h:5-119).  Every function writes in place like the reference and returns
None; argument errors raise RuntimeError (the reference's TORCH_CHECK)."""
from typing import Optional

import torch

from ._common import check, device_of, float_code, lib, require_cuda, stream_of


def _kv_fp8(kv_cache_dtype: str) -> int:
    if kv_cache_dtype == "auto":
        return 0
    if kv_cache_dtype == "fp8":
        return 1
    raise RuntimeError(f"Unsupported data type of kv cache: {kv_cache_dtype}")


def _attn_common(query, key_cache, value_cache, block_tables, context_lens, kv_cache_dtype):
    require_cuda(query, "query")
    code = float_code(query, "query")
    fp8 = _kv_fp8(kv_cache_dtype)
    if fp8:
        if key_cache.dtype != torch.uint8:
            raise RuntimeError("kv_cache_dtype 'fp8' needs a uint8 KV cache")
    elif key_cache.dtype != query.dtype:
        raise RuntimeError(
            f"kv cache dtype {key_cache.dtype} does not match query dtype {query.dtype}")
    if block_tables.dtype != torch.int32 or context_lens.dtype != torch.int32:
        raise RuntimeError("block_tables and context_lens must be int32")
    if query.stride(2) != 1 or query.stride(1) != query.size(2):
        raise RuntimeError("query must be contiguous in its last two dimensions")
    if key_cache.stride(1) != value_cache.stride(1) or key_cache.stride(0) != value_cache.stride(0):
        raise RuntimeError("key_cache and value_cache must share block/head strides")
    return code, fp8


def paged_attention_v1(out: torch.Tensor, query: torch.Tensor, key_cache: torch.Tensor,
                       value_cache: torch.Tensor, num_kv_heads: int, scale: float,
                       block_tables: torch.Tensor, context_lens: torch.Tensor, block_size: int,
                       max_context_len: int, alibi_slopes: Optional[torch.Tensor],
                       kv_cache_dtype: str, kv_scale: float) -> None:
    """csrc/attention/attention_kernels.cu:741-778."""
    code, fp8 = _attn_common(query, key_cache, value_cache, block_tables, context_lens,
                             kv_cache_dtype)
    num_seqs, num_heads, head_size = query.shape
    with device_of(query):
        check(lib.vltr_paged_attention_v1(
            out.data_ptr(), query.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
            num_seqs, num_heads, head_size, num_kv_heads, scale, block_tables.data_ptr(),
            context_lens.data_ptr(), block_size, max_context_len, block_tables.size(1),
            alibi_slopes.data_ptr() if alibi_slopes is not None else None, query.stride(0),
            key_cache.stride(0), key_cache.stride(1), code, fp8, kv_scale, stream_of(query)))


def paged_attention_v2(out: torch.Tensor, exp_sums: torch.Tensor, max_logits: torch.Tensor,
                       tmp_out: torch.Tensor, query: torch.Tensor, key_cache: torch.Tensor,
                       value_cache: torch.Tensor, num_kv_heads: int, scale: float,
                       block_tables: torch.Tensor, context_lens: torch.Tensor, block_size: int,
                       max_context_len: int, alibi_slopes: Optional[torch.Tensor],
                       kv_cache_dtype: str, kv_scale: float) -> None:
    """csrc/attention/attention_kernels.cu:935-975."""
    code, fp8 = _attn_common(query, key_cache, value_cache, block_tables, context_lens,
                             kv_cache_dtype)
    num_seqs, num_heads, head_size = query.shape
    need = (max_context_len + 511) // 512
    if exp_sums.dim() != 3 or exp_sums.size(2) < need or tmp_out.size(2) < need:
        raise RuntimeError("exp_sums/max_logits/tmp_out too small for max_context_len")
    if exp_sums.size(2) != need:
        raise RuntimeError("exp_sums/max_logits/tmp_out must have ceil(max_context_len/512) partitions")
    with device_of(query):
        check(lib.vltr_paged_attention_v2(
            out.data_ptr(), exp_sums.data_ptr(), max_logits.data_ptr(), tmp_out.data_ptr(),
            query.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(), num_seqs, num_heads,
            head_size, num_kv_heads, scale, block_tables.data_ptr(), context_lens.data_ptr(),
            block_size, max_context_len, block_tables.size(1),
            alibi_slopes.data_ptr() if alibi_slopes is not None else None, query.stride(0),
            key_cache.stride(0), key_cache.stride(1), code, fp8, kv_scale, stream_of(query)))


# ---- activations (csrc/activation_kernels.cu:76-161) ------------------------------------------
def _act(out: torch.Tensor, input: torch.Tensor, act: int, gated: bool) -> None:
    require_cuda(input, "input")
    code = float_code(input, "input")
    last = input.size(-1)
    d = last // 2 if gated else last
    num_tokens = input.numel() // last if last else 0
    with device_of(input):
        check(lib.vltr_activation(out.data_ptr(), input.data_ptr(), num_tokens, d, act, code,
                                  stream_of(input)))


def silu_and_mul(out: torch.Tensor, input: torch.Tensor) -> None:
    _act(out, input, 0, True)


def gelu_and_mul(out: torch.Tensor, input: torch.Tensor) -> None:
    _act(out, input, 1, True)


def gelu_tanh_and_mul(out: torch.Tensor, input: torch.Tensor) -> None:
    _act(out, input, 2, True)


def gelu_new(out: torch.Tensor, input: torch.Tensor) -> None:
    _act(out, input, 3, False)


def gelu_fast(out: torch.Tensor, input: torch.Tensor) -> None:
    _act(out, input, 4, False)


# ---- layernorm (csrc/layernorm_kernels.cu:277-352) ----------------------------------------------
def rms_norm(out: torch.Tensor, input: torch.Tensor, weight: torch.Tensor,
             epsilon: float) -> None:
    require_cuda(input, "input")
    code = float_code(input, "input")
    hidden = input.size(-1)
    num_tokens = input.numel() // hidden if hidden else 0
    with device_of(input):
        check(lib.vltr_rms_norm(out.data_ptr(), input.data_ptr(), weight.data_ptr(), epsilon,
                                num_tokens, hidden, code, stream_of(input)))


def fused_add_rms_norm(input: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor,
                       epsilon: float) -> None:
    require_cuda(input, "input")
    code = float_code(input, "input")
    hidden = input.size(-1)
    num_tokens = input.numel() // hidden if hidden else 0
    with device_of(input):
        check(lib.vltr_fused_add_rms_norm(input.data_ptr(), residual.data_ptr(),
                                          weight.data_ptr(), epsilon, num_tokens, hidden, code,
                                          stream_of(input)))


# ---- rotary (csrc/pos_encoding_kernels.cu:124-226) ----------------------------------------------
def rotary_embedding(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                     head_size: int, cos_sin_cache: torch.Tensor, is_neox: bool) -> None:
    require_cuda(query, "query")
    code = float_code(query, "query")
    if positions.dtype != torch.int64:
        raise RuntimeError("positions must be int64")
    num_tokens = query.numel() // query.size(-1)
    with device_of(query):
        check(lib.vltr_rotary_embedding(
            positions.data_ptr(), query.data_ptr(), key.data_ptr(), num_tokens,
            query.size(-1) // head_size, key.size(-1) // head_size, head_size,
            cos_sin_cache.size(1), query.stride(-2), key.stride(-2), cos_sin_cache.data_ptr(),
            None, 1 if is_neox else 0, code, stream_of(query)))


def batched_rotary_embedding(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                             head_size: int, cos_sin_cache: torch.Tensor, is_neox: bool,
                             rot_dim: int, cos_sin_cache_offsets: torch.Tensor) -> None:
    require_cuda(query, "query")
    code = float_code(query, "query")
    if positions.dtype != torch.int64 or cos_sin_cache_offsets.dtype != torch.int64:
        raise RuntimeError("positions and cos_sin_cache_offsets must be int64")
    num_tokens = cos_sin_cache_offsets.size(0)
    with device_of(query):
        check(lib.vltr_rotary_embedding(
            positions.data_ptr(), query.data_ptr(), key.data_ptr(), num_tokens,
            query.size(-1) // head_size, key.size(-1) // head_size, head_size, rot_dim,
            query.stride(-2), key.stride(-2), cos_sin_cache.data_ptr(),
            cos_sin_cache_offsets.data_ptr(), 1 if is_neox else 0, code, stream_of(query)))
