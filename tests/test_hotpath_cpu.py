"""CPU: the host-side call sequence of tools/hotpath.py DecodeHotPath driven through the
oracle dressed as a vllm._C module (bench.OraclePort) - the same wiring the benchmark's reference
arm uses with the reference's CPU kernels.  Checks the plumbing (buffers, views, optional
projections), not kernels."""
import numpy as np
import torch

from bench import OraclePort, attention_bytes, build_inputs, step_bytes, trace_contexts
from hotpath import SHAPES, DecodeHotPath, ModelShape


def _tiny():
    s = ModelShape(**{**SHAPES["llama-3-8b"].__dict__})
    s.hidden, s.q_heads, s.kv_heads, s.head_size, s.inter = 256, 4, 2, 64, 512
    return s


def test_step_writes_cache_and_is_deterministic():
    s = _tiny()
    ctx = np.array([1, 16, 17, 100])
    inp = build_inputs(ctx, 16, 64, 0)
    h = (torch.randn(4, s.hidden, generator=torch.Generator().manual_seed(0)) * 0.5).to(s.dtype)
    outs = []
    for _ in range(2):
        hp = DecodeHotPath(s, 64, 4, "cpu", layers=2, C=OraclePort())
        before = [c.clone() for c in hp.kv_caches]
        out = hp.step(h, inp["positions"], inp["slot_mapping"], inp["block_tables"],
                      inp["context_lens"], int(ctx.max()))
        assert out.shape == (4, s.hidden) and torch.isfinite(out.float()).all()
        # every sequence's newest token landed in its slot of every layer's cache
        for b, a in zip(before, hp.kv_caches):
            assert not torch.equal(b, a)
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1])


def test_step_with_projections_uses_the_weights():
    s = _tiny()
    ctx = np.array([5, 33])
    inp = build_inputs(ctx, 16, 32, 0)
    h = (torch.randn(2, s.hidden, generator=torch.Generator().manual_seed(1)) * 0.5).to(s.dtype)
    res = []
    for seed in (1, 2):
        hp = DecodeHotPath(s, 32, 2, "cpu", layers=2, C=OraclePort())
        hp.enable_gemms(seed=seed)
        assert len(hp.gemm_w) == 2 and hp.gemm_w[0][0].shape == ((4 + 2 * 2) * 64, 256)
        res.append(hp.step(h, inp["positions"], inp["slot_mapping"], inp["block_tables"],
                           inp["context_lens"], int(ctx.max())).clone())
        assert torch.isfinite(res[-1].float()).all()
    assert not torch.equal(res[0], res[1])       # the projections are really in the data path


def test_algorithmic_byte_formulas():
    """KV once per kv head, q + out, block-table entries, context lengths."""
    s = SHAPES["llama-3-8b"]
    ctx = np.array([16, 17, 1000])
    kv = (16 + 17 + 1000) * 2 * 8 * 128 * 2
    qo = 2 * 3 * 32 * 128 * 2
    tbl = 4 * (1 + 2 + 63) + 4 * 3
    assert attention_bytes(ctx, s, 16) == kv + qo + tbl
    assert step_bytes(ctx, s, 16, 2) > 2 * attention_bytes(ctx, s, 16)
    c = trace_contexts(256)
    assert len(c) == 256 and c.min() >= 1 and c.max() <= 8192 and int(c.sum()) == 103012


def test_tensor_parallel_shapes():
    """Megatron split of the reference (llama.py:96-111): q heads and MLP width divided, kv heads
    divided or replicated."""
    s70 = SHAPES["llama-2-70b"].shard(8)
    assert (s70.q_heads, s70.kv_heads, s70.inter, s70.hidden) == (8, 1, 3584, 8192)
    s7 = SHAPES["llama-2-7b"].shard(8)
    assert (s7.q_heads, s7.kv_heads, s7.inter) == (4, 4, 1376)
    s8 = SHAPES["llama-3-8b"].shard(4)
    assert (s8.q_heads, s8.kv_heads, s8.inter) == (8, 2, 3584)
    assert SHAPES["llama-3-8b"].shard(1) is SHAPES["llama-3-8b"]
    s16 = ModelShape("x", 1024, 2, 16, 2, 64, 2048, torch.bfloat16).shard(8)
    assert (s16.q_heads, s16.kv_heads) == (2, 1)          # kv heads replicated below one per rank
