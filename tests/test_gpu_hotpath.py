"""The decode step as a whole (tools/hotpath.py: the per-layer op sequence of
llama.py:215-238 through the vllm._C mirror): against the oracle driven through the same sequence,
and the launch modes (programmatic dependent launch on/off, CUDA-graph replay) against each other,
which must agree bit for bit - they run the same kernels on the same data."""
import numpy as np
import pytest
import torch

from bench import OraclePort, build_inputs, trace_contexts
from hotpath import SHAPES, DecodeHotPath, ModelShape

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _small(name, dtype):
    s = ModelShape(**{**SHAPES[name].__dict__})
    s.dtype = dtype
    return s


def _cpu_twin(hp: DecodeHotPath, shape, num_blocks, T, layers):
    """A CPU instance with the SAME state (KV pool, weights, stand-in GEMM outputs), driven by
    the oracle dressed as a vllm._C module."""
    tw = DecodeHotPath(shape, num_blocks, T, "cpu", layers=layers, C=OraclePort())
    for name in ("qkv", "o_proj_out", "gate_up", "down_out", "hidden", "cos_sin"):
        getattr(tw, name).copy_(getattr(hp, name).cpu())
    for a, b in zip(tw.kv_caches, hp.kv_caches):
        a.copy_(b.cpu())
    for a, b in zip(tw.norm_w, hp.norm_w):
        a.copy_(b.cpu())
    return tw


@pytest.mark.parametrize("name,dtype", [("llama-2-7b", torch.float16), ("llama-3-8b", torch.bfloat16)])
def test_step_matches_oracle(name, dtype):
    shape = _small(name, dtype)
    ctx = np.array([1, 15, 16, 17, 100, 333, 512, 700], dtype=np.int64)
    B, layers, nb = len(ctx), 2, 256
    inp = build_inputs(ctx, 16, nb, 0)
    hp = DecodeHotPath(shape, nb, B, DEV, layers=layers)
    tw = _cpu_twin(hp, shape, nb, B, layers)
    h = (torch.randn(B, shape.hidden) * 0.5).to(dtype)
    got = hp.step(h.to(DEV), *(inp[k].to(DEV) for k in ("positions", "slot_mapping", "block_tables",
                                                         "context_lens")), int(ctx.max()))
    want = tw.step(h, inp["positions"], inp["slot_mapping"], inp["block_tables"],
                   inp["context_lens"], int(ctx.max()))
    # the step's result is the final fused_add_rms_norm of stand-in GEMM outputs and the residual
    # chain: element-wise ops in T, bit-exact per op (tests/test_gpu_elementwise.py) up to the
    # fp32 association order of the row sums -> 1 ulp of T on O(1) values
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3
    assert (got.cpu().float() - want.float()).abs().max().item() <= tol
    # the caches now hold the rotated keys / values of the new tokens: exact copies
    for a, b in zip(tw.kv_caches, hp.kv_caches):
        ka, kb = a.float(), b.cpu().float()
        assert (ka - kb).abs().max().item() <= tol
    # attention output of the last layer against the oracle
    assert (hp.attn_out[:B].cpu().float() - tw.attn_out[:B].float()).abs().max().item() <= 1e-2


@pytest.mark.parametrize("name", ["llama-2-7b", "llama-3-8b", "llama-3-8b-fp8kv"])
def test_launch_modes_agree_bitwise(name):
    from vllm_ltr_b200._lib import tune
    shape = _small(name, torch.bfloat16)
    ctx = trace_contexts(64)
    nb = int(((ctx + 15) // 16).sum()) + 16
    inp = {k: v.to(DEV) for k, v in build_inputs(ctx, 16, nb, 0).items()}
    h = (torch.randn(len(ctx), shape.hidden, device=DEV) * 0.5).to(shape.dtype)
    outs = {}
    names = ("qkv", "o_proj_out", "gate_up", "down_out")

    def snapshot(hp):
        return [getattr(hp, n).clone() for n in names] + [c.clone() for c in hp.kv_caches]

    def restore(hp, snap):
        # a step rewrites the stand-in GEMM outputs in place (rotary on qkv, the fused norm on the
        # row-parallel outputs) and appends to the caches
        for n, t in zip(names, snap):
            getattr(hp, n).copy_(t)
        for c, t in zip(hp.kv_caches, snap[len(names):]):
            c.copy_(t)

    try:
        for mode in ("pdl", "plain", "graph"):
            tune("pdl", 0 if mode == "plain" else 1)
            hp = DecodeHotPath(shape, nb, len(ctx), DEV, layers=3, seed=1)
            snap = snapshot(hp)
            args = (h, inp["positions"], inp["slot_mapping"], inp["block_tables"],
                    inp["context_lens"], int(ctx.max()))
            if mode == "graph":
                g, out = hp.capture(*args)     # runs one uncaptured step first
                restore(hp, snap)
                g.replay()
                first = out.clone()
                restore(hp, snap)
                g.replay()
                torch.cuda.synchronize()
                assert torch.equal(first, out), "graph replay is not repeatable"
            else:
                out = hp.step(*args)
            torch.cuda.synchronize()
            outs[mode] = (out.clone(), hp.attn_out.clone(), [c.clone() for c in hp.kv_caches])
    finally:
        tune("pdl", 1)
    for mode in ("plain", "graph"):
        assert torch.equal(outs["pdl"][0], outs[mode][0]), mode
        assert torch.equal(outs["pdl"][1], outs[mode][1]), mode
        for a, b in zip(outs["pdl"][2], outs[mode][2]):
            assert torch.equal(a, b), mode
