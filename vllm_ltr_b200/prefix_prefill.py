"""Prefill attention over the paged KV cache - drop-in for the reference's Triton kernel
``vllm.attention.ops.prefix_prefill.context_attention_fwd`` (prefix_prefill.py:630-755), which
``PagedAttention.forward_prefix`` (vllm/attention/ops/paged_attn.py:163-195) calls for every
prefill chunk when chunked prefill is enabled.  Same signature, same in-place ``o``; the work is
done by ``vltr_context_attention_fwd`` (csrc/prefill_attention.cu).

``install()`` rebinds the name inside the reference's ``paged_attn`` module, so the unmodified
engine stops importing Triton on the step path.
"""
from typing import Optional

import torch

from ._C._common import DTYPE_CODE, check, device_of, lib, require_cuda, stream_of
from ._lib import declare
import ctypes

_P = ctypes.c_void_p
declare("vltr_context_attention_fwd", ctypes.c_int,
        [_P] * 10 + [ctypes.c_int] * 7 + [ctypes.c_float, _P] + [ctypes.c_int64] * 6 +
        [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P])


@torch.inference_mode()
def context_attention_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor,
                          k_cache: torch.Tensor, v_cache: torch.Tensor, b_loc: torch.Tensor,
                          b_start_loc: torch.Tensor, b_seq_len: torch.Tensor,
                          b_ctx_len: torch.Tensor, max_input_len: int,
                          alibi_slopes: Optional[torch.Tensor] = None,
                          cache_has_new_tokens: bool = False) -> None:
    """q, o: [num_tokens, num_heads, head_size]; k, v: [num_tokens, num_kv_heads, head_size];
    k_cache [blocks, kv_heads, head_size/x, block_size, x], v_cache [blocks, kv_heads, head_size,
    block_size]; b_loc: block tables [batch, max_blocks]; b_start_loc / b_seq_len / b_ctx_len:
    [batch] int32.  ``cache_has_new_tokens=True`` skips re-writing the chunk's k/v into the cache
    (the engine has already done it, flash_attn.py:192)."""
    require_cuda(q, "q")
    code = DTYPE_CODE.get(q.dtype)
    if q.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError(f"Unsupported data type: {q.dtype} (context_attention_fwd: f16 / bf16)")
    fp8 = 1 if k_cache.dtype == torch.uint8 else 0
    if not fp8 and k_cache.dtype != q.dtype:
        raise RuntimeError("kv cache dtype does not match query dtype")
    # the engine passes int32 metadata, the reference's unit test int64: both are accepted
    def i32(t, name):
        if t.dtype == torch.int32:
            return t if t.is_contiguous() else t.contiguous()
        if t.dtype == torch.int64:
            return t.to(torch.int32).contiguous()
        raise RuntimeError(f"{name} must be an int32 or int64 tensor")
    b_loc, b_start_loc = i32(b_loc, "b_loc"), i32(b_start_loc, "b_start_loc")
    b_seq_len, b_ctx_len = i32(b_seq_len, "b_seq_len"), i32(b_ctx_len, "b_ctx_len")
    if q.stride(2) != 1 or q.stride(1) != q.size(2) or o.stride(2) != 1 or o.stride(1) != o.size(2):
        raise RuntimeError("q / o must be contiguous in their last two dimensions")
    num_tokens, num_heads, head_size = q.shape
    num_kv_heads = k.shape[1]
    with device_of(q):
        check(lib.vltr_context_attention_fwd(
            o.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), k_cache.data_ptr(),
            v_cache.data_ptr(), b_loc.data_ptr(), b_start_loc.data_ptr(), b_seq_len.data_ptr(),
            b_ctx_len.data_ptr(), b_seq_len.shape[0], num_tokens, num_heads, num_kv_heads,
            head_size, v_cache.shape[3], b_loc.shape[1], 1.0 / (head_size ** 0.5),
            alibi_slopes.data_ptr() if alibi_slopes is not None else None, q.stride(0),
            k.stride(0), v.stride(0), o.stride(0), k_cache.stride(0), k_cache.stride(1), code, fp8,
            1 if cache_has_new_tokens else 0, stream_of(q)))


def install(engine_path: bool = True) -> None:
    """Point the reference's PagedAttention.forward_prefix at this implementation.

    ``engine_path=True`` (the default): the only caller is ``FlashAttentionImpl.forward``, which
    has already stored the chunk's k / v with ``write_to_paged_cache`` (flash_attn.py:192) before
    it reaches ``forward_prefix`` (:240-252) - so the kernel is bound with
    ``cache_has_new_tokens=True`` and does not scatter them a second time (one launch and one
    pass over the chunk's cache lines less per layer).  ``engine_path=False`` keeps the
    stand-alone semantics of the Triton kernel (cache holds only the prefix; used by the unit
    tests, which call it the way tests/kernels/test_prefix_prefill.py does).

    Numerical note for an fp8 (e5m2) cache: the chunk's own keys / values are read back from the
    cache, i.e. in their e5m2-rounded form, where the Triton kernel uses the unquantised k / v
    of the current chunk (and e5m2 values for the prefix).  Decode attention sees the quantised
    values either way."""
    import functools

    import vllm.attention.ops.paged_attn as pa
    pa.context_attention_fwd = (functools.partial(context_attention_fwd, cache_has_new_tokens=True)
                                if engine_path else context_attention_fwd)
