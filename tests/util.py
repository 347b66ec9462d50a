"""Shared seeded input builders for the parity tests (mirrors the fixtures of the reference's
tests/kernels/*.py: tests/kernels/conftest.py -> vllm/utils.py:344-410 create_kv_caches_with_random,
test_attention.py:140-166 block tables / context lengths)."""
import random
from typing import List, Optional, Tuple

import torch

from oracle import kernels as K

NUM_BLOCKS = 4321        # tests/kernels/test_attention.py:20
PARTITION_SIZE = 512     # tests/kernels/test_attention.py:21


def elem_size(dtype: torch.dtype) -> int:
    return torch.tensor([], dtype=dtype).element_size()


def make_cache(num_blocks, block_size, num_kv_heads, head_size, dtype, kv_cache_dtype, gen, device):
    """Random paged caches, uniform(-1/sqrt(D), 1/sqrt(D)) (vllm/utils.py:344-410), generated on
    `device`.  fp8: the same values rounded to e5m2 bytes by the oracle's converter."""
    scale = head_size ** -0.5
    x = 16 // (1 if kv_cache_dtype == "fp8" else elem_size(dtype))
    kshape = (num_blocks, num_kv_heads, head_size // x, block_size, x)
    vshape = (num_blocks, num_kv_heads, head_size, block_size)
    k = ((torch.rand(kshape, generator=gen, device=device) * 2 - 1) * scale).to(dtype)
    v = ((torch.rand(vshape, generator=gen, device=device) * 2 - 1) * scale).to(dtype)
    if kv_cache_dtype == "fp8":
        return K.fp8_e5m2_from_fast(k), K.fp8_e5m2_from_fast(v)
    return k, v


def attention_case(num_seqs: int, num_heads: Tuple[int, int], head_size: int, block_size: int,
                   dtype: torch.dtype, kv_cache_dtype: str, use_alibi: bool, seed: int,
                   max_len: int, device: str, context_lens: Optional[List[int]] = None,
                   num_blocks: int = NUM_BLOCKS):
    """Inputs of one paged-attention call from seeded generators on `device` (the same tensors
    feed the oracle and the kernel)."""
    rnd = random.Random(seed)
    g = torch.Generator(device=device).manual_seed(seed)
    scale = float(1.0 / (head_size ** 0.5))
    hq, hkv = num_heads
    query = ((torch.rand(num_seqs, hq, head_size, generator=g, device=device) * 2 - 1)
             * scale).to(dtype)
    alibi = (torch.randn(hq, generator=g, dtype=torch.float32, device=device)
             if use_alibi else None)
    if context_lens is None:
        context_lens = [rnd.randint(1, max_len) for _ in range(num_seqs)]
        context_lens[-1] = max_len
    max_ctx = max(context_lens) if context_lens else 0
    mb = max(1, (max_ctx + block_size - 1) // block_size)
    bt = torch.tensor([[rnd.randint(0, num_blocks - 1) for _ in range(mb)]
                       for _ in range(num_seqs)], dtype=torch.int32).reshape(num_seqs, mb)
    kc, vc = make_cache(num_blocks, block_size, hkv, head_size, dtype, kv_cache_dtype, g, device)
    return dict(query=query, key_cache=kc, value_cache=vc, block_tables=bt.to(device),
                context_lens=torch.tensor(context_lens, dtype=torch.int32).to(device),
                alibi=alibi, scale=scale, max_ctx=max_ctx, num_kv_heads=hkv,
                block_size=block_size, kv_cache_dtype=kv_cache_dtype)


def run_attention(ops, c, version: str, out_dtype=None):
    """Call ops.paged_attention_v1/v2 the way vllm/attention/ops/paged_attn.py:112-161 does."""
    q = c["query"]
    out = torch.empty_like(q)
    side = None
    if version == "v1":
        ops.paged_attention_v1(out, q, c["key_cache"], c["value_cache"], c["num_kv_heads"],
                               c["scale"], c["block_tables"], c["context_lens"], c["block_size"],
                               c["max_ctx"], c["alibi"], c["kv_cache_dtype"], 1.0)
    else:
        B, H, D = q.shape
        P = (c["max_ctx"] + PARTITION_SIZE - 1) // PARTITION_SIZE
        tmp = torch.empty(B, H, P, D, dtype=q.dtype, device=q.device)
        es = torch.empty(B, H, P, dtype=torch.float32, device=q.device)
        ml = torch.empty_like(es)
        ops.paged_attention_v2(out, es, ml, tmp, q, c["key_cache"], c["value_cache"],
                               c["num_kv_heads"], c["scale"], c["block_tables"],
                               c["context_lens"], c["block_size"], c["max_ctx"], c["alibi"],
                               c["kv_cache_dtype"], 1.0)
        side = (ml, es, tmp)
    return out, side


def oracle_attention(c):
    return K.paged_attention(c["query"], c["key_cache"], c["value_cache"], c["block_tables"],
                             c["context_lens"], c["scale"], c["alibi"], c["kv_cache_dtype"])


def cos_sin_cache(head_size: int, rot_dim: int, max_pos: int, dtype: torch.dtype,
                  base: float = 10000.0) -> torch.Tensor:
    """vllm/model_executor/layers/rotary_embedding.py:70-93 (_compute_cos_sin_cache)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, rot_dim, 2, dtype=torch.float) / rot_dim))
    t = torch.arange(max_pos, dtype=torch.float)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(dtype)


def ulp_diff(a: torch.Tensor, b: torch.Tensor) -> int:
    """max distance in representable values between two same-dtype f16/bf16 tensors."""
    assert a.dtype == b.dtype and a.dtype in (torch.float16, torch.bfloat16)
    ia = a.contiguous().view(torch.int16).to(torch.int32)
    ib = b.contiguous().view(torch.int16).to(torch.int32)
    # map sign-magnitude to a monotone integer line
    ia = torch.where(ia < 0, -(ia & 0x7FFF), ia)
    ib = torch.where(ib < 0, -(ib & 0x7FFF), ib)
    return int((ia - ib).abs().max()) if a.numel() else 0
