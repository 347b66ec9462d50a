"""GPU parity against tests/golden/ref_cuda.npz: outputs of the reference's own CUDA kernels
(recompiled for sm_100a, run on a B200 by tests/golden/gen_golden_cuda.py).  f16/bf16 element-wise
and index/copy ops: bit-exact.  Reductions (rms_norm) within one rounding step; attention inside
the reference test's tolerance of the reference kernel's output."""
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "ref_cuda.npz")
DT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(PATH):
        pytest.skip("tests/golden/ref_cuda.npz not generated yet (needs a GPU box)")
    return np.load(PATH)


@pytest.fixture(scope="module")
def C():
    from vllm_ltr_b200 import _C
    return _C


def t(a, dt):
    return torch.from_numpy(np.asarray(a)).to(dt).to(DEV)


def same(a, b):
    if a.dtype == torch.float32:
        return torch.allclose(a, b, atol=1e-6, rtol=1e-6)
    return torch.equal(a, b)


@pytest.mark.parametrize("name", list(DT))
def test_golden_activations(gold, C, name):
    dt = DT[name]
    x = t(gold[f"act.{name}.x"], dt)
    for act in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
        o = torch.empty(5, 128, dtype=dt, device=DEV)
        getattr(C.ops, act)(o, x)
        assert same(o, t(gold[f"{act}.{name}"], dt)), act
    for act in ("gelu_new", "gelu_fast"):
        o = torch.empty_like(x)
        getattr(C.ops, act)(o, x)
        assert same(o, t(gold[f"{act}.{name}"], dt)), act


@pytest.mark.parametrize("name", list(DT))
def test_golden_rms(gold, C, name):
    dt = DT[name]
    for tag, n in (("", 512), ("771", 771)):
        x, r, w = (t(gold[f"{'rms' if not tag else 'frms771'}.{name}.{k}"], dt) for k in "xrw")
        if not tag:
            o = torch.empty_like(x)
            C.ops.rms_norm(o, x, w, 1e-5)
            ref = t(gold[f"rms.{name}.out"], dt)
            assert torch.allclose(o, ref, atol=1e-5, rtol=1e-5) if dt == torch.float32 \
                else util.ulp_diff(o, ref) <= 2
        C.ops.fused_add_rms_norm(x, r, w, 1e-5)
        rx, rr = t(gold[f"frms{tag}.{name}.out_x"], dt), t(gold[f"frms{tag}.{name}.out_r"], dt)
        assert same(r, rr)
        assert torch.allclose(x, rx, atol=1e-5, rtol=1e-5) if dt == torch.float32 \
            else util.ulp_diff(x, rx) <= 2


@pytest.mark.parametrize("name", list(DT))
@pytest.mark.parametrize("neox", [1, 0])
def test_golden_rope(gold, C, name, neox):
    dt = DT[name]
    q, k = t(gold[f"rope.{name}.q"], dt), t(gold[f"rope.{name}.k"], dt)
    C.ops.rotary_embedding(torch.from_numpy(gold[f"rope.{name}.pos"]).to(DEV), q, k, 64,
                           t(gold[f"rope.{name}.cache"], dt), bool(neox))
    assert same(q, t(gold[f"rope.{name}.neox{neox}.q"], dt))
    assert same(k, t(gold[f"rope.{name}.neox{neox}.k"], dt))


@pytest.mark.parametrize("name", list(DT))
@pytest.mark.parametrize("kv", ["auto", "fp8"])
@pytest.mark.parametrize("bs", [16, 32])
def test_golden_attention_and_cache(gold, C, name, kv, bs):
    dt = DT[name]
    cdt = torch.uint8 if kv == "fp8" else dt
    tag = f"attn.{name}.{kv}.bs{bs}"
    c = dict(query=t(gold[f"{tag}.q"], dt), key_cache=t(gold[f"{tag}.kc"], cdt),
             value_cache=t(gold[f"{tag}.vc"], cdt),
             block_tables=torch.from_numpy(gold[f"{tag}.bt"]).to(DEV),
             context_lens=torch.from_numpy(gold[f"{tag}.cl"]).to(DEV),
             alibi=t(gold[f"{tag}.alibi"], torch.float32), scale=64 ** -0.5, max_ctx=700,
             num_kv_heads=2, block_size=bs, kv_cache_dtype=kv)
    atol = 1e-2 if kv == "fp8" else 1e-3
    for ver in ("v1", "v2"):
        out, _ = util.run_attention(C.ops, c, ver)
        ref = t(gold[f"{tag}.{ver}"], torch.float32)
        assert torch.allclose(out.float(), ref, atol=atol, rtol=1e-5), \
            (ver, (out.float() - ref).abs().max().item())
    kc, vc = c["key_cache"].clone(), c["value_cache"].clone()
    C.cache_ops.reshape_and_cache(t(gold[f"{tag}.key"], dt), t(gold[f"{tag}.val"], dt), kc, vc,
                                  torch.from_numpy(gold[f"{tag}.slots"]).to(DEV), kv, 1.0)
    assert torch.equal(kc, t(gold[f"{tag}.kc_out"], cdt))
    assert torch.equal(vc, t(gold[f"{tag}.vc_out"], cdt))


@pytest.mark.parametrize("name", list(DT))
def test_golden_convert_fp8(gold, C, name):
    dt = DT[name]
    src = t(gold[f"fp8.{name}.src"], dt)
    dst = torch.empty(src.shape, dtype=torch.uint8, device=DEV)
    C.cache_ops.convert_fp8(src, dst)
    assert torch.equal(dst.cpu(), torch.from_numpy(gold[f"fp8.{name}.codes"]))
    back = torch.empty(256, dtype=dt, device=DEV)
    C.cache_ops.convert_fp8(torch.arange(256, dtype=torch.uint8, device=DEV), back)
    ref = torch.from_numpy(gold[f"fp8.{name}.back"])
    fin = torch.isfinite(ref)
    assert torch.equal(back.cpu().float()[fin], ref[fin])
