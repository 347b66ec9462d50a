"""Pins the oracle (oracle/kernels.py) against the reference itself - CPU only.

1. tests/golden/ref_cpu.npz: outputs of the reference's C++ CPU kernels (csrc/cpu/*.cpp, compiled
   unmodified into oracle/_ref, run by tests/golden/gen_golden.py).  Those kernels keep fp32
   intermediates where the CUDA kernels round to T (e.g. silu(x)*y is rounded once on CPU, twice
   on CUDA), so bf16 results agree to one rounding step, f32 to float tolerance; index/copy ops
   are exact.
2. tests/golden/ref_py.npz: outputs of the reference's torch formulas (RMSNorm._forward,
   RotaryEmbedding._forward, SiluAndMul._forward ... imported from /root/reference) - the oracles
   of the reference's own kernel tests, compared at those tests' tolerances.
3. the live oracle/_ref/ref_cpu_C module when present (fresh seeds, larger shapes).
4. fp8 e5m2: the numpy restatement vs torch.float8_e5m2 over all 65536 half bit patterns.
"""
import os

import numpy as np
import pytest
import torch

from oracle import kernels as K
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
DT = {"bfloat16": torch.bfloat16, "float32": torch.float32, "float16": torch.float16}
TOL = {torch.half: (1e-3, 1e-3), torch.bfloat16: (1e-3, 1.6e-2), torch.float: (1e-5, 1.3e-6)}


@pytest.fixture(scope="module")
def gc():
    return np.load(os.path.join(HERE, "golden", "ref_cpu.npz"))


@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(HERE, "golden", "ref_py.npz"))


def t(a, dt):
    return torch.from_numpy(np.asarray(a)).to(dt)


def close_T(out, ref, dt, steps=1):
    """equal up to `steps` representable values (bf16/f16) or float tolerance (f32)."""
    if dt == torch.float32:
        assert torch.allclose(out, ref, atol=2e-6, rtol=2e-6)
    else:
        assert util.ulp_diff(out, ref) <= steps


# ---- 1. golden vectors from the reference's C++ CPU kernels ------------------------------------
@pytest.mark.parametrize("name", ["bfloat16", "float32"])
@pytest.mark.parametrize("act", ["silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul", "gelu_new",
                                 "gelu_fast"])
def test_golden_cpu_activation(gc, name, act):
    dt = DT[name]
    out = K.activation(t(gc[f"{act}.{name}.in"], dt), act)
    ref = t(gc[f"{act}.{name}.out"], dt)
    if act.endswith("_and_mul") or dt == torch.float32:
        close_T(out, ref, dt)
    else:
        # gelu_new / gelu_fast: the CUDA kernels round every intermediate to T
        # (activation_kernels.cu:132-144), the CPU kernels keep fp32 - 1 + tanh(.) cancels for
        # negative x, so only an absolute band of one bf16 step at O(1) is meaningful here
        assert torch.allclose(out.float(), ref.float(), atol=8e-3, rtol=2e-2)


@pytest.mark.parametrize("name", ["bfloat16", "float32"])
def test_golden_cpu_rms(gc, name):
    dt = DT[name]
    x, w = t(gc[f"rms.{name}.x"], dt), t(gc[f"rms.{name}.w"], dt)
    close_T(K.rms_norm(x, w, 1e-5), t(gc[f"rms.{name}.out"], dt), dt)
    ox, orr = K.fused_add_rms_norm(x, t(gc[f"frms.{name}.r"], dt), w, 1e-5)
    assert torch.equal(orr, t(gc[f"frms.{name}.out_r"], dt))
    close_T(ox, t(gc[f"frms.{name}.out_x"], dt), dt)


@pytest.mark.parametrize("name", ["bfloat16", "float32"])
@pytest.mark.parametrize("neox", [True, False])
def test_golden_cpu_rope(gc, name, neox):
    dt = DT[name]
    q, k = t(gc[f"rope.{name}.q"], dt), t(gc[f"rope.{name}.k"], dt)
    oq, ok = K.rotary_embedding(torch.from_numpy(gc[f"rope.{name}.pos"]), q, k, 64,
                                t(gc[f"rope.{name}.cache"], dt), neox)
    # CUDA rounds x*cos and y*sin to T before the subtraction (pos_encoding_kernels.cu:34-38), the
    # CPU kernel rounds once at the end: the two differ by at most the two product roundings,
    # eps(T) * max|input|, plus the final rounding
    eps = {torch.bfloat16: 2.0 ** -8, torch.float32: 2.0 ** -23}[dt]
    atol = 2 * eps * float(max(q.abs().max(), k.abs().max()))
    assert torch.allclose(oq, t(gc[f"rope.{name}.neox{int(neox)}.q"], dt), atol=atol, rtol=2 * eps)
    assert torch.allclose(ok, t(gc[f"rope.{name}.neox{int(neox)}.k"], dt), atol=atol, rtol=2 * eps)


@pytest.mark.parametrize("name", ["bfloat16", "float32"])
def test_golden_cpu_cache_ops_exact(gc, name):
    dt = DT[name]
    kc, vc = t(gc[f"cache.{name}.kc"], dt), t(gc[f"cache.{name}.vc"], dt)
    K.reshape_and_cache(t(gc[f"cache.{name}.key"], dt), t(gc[f"cache.{name}.val"], dt), kc, vc,
                        torch.from_numpy(gc[f"cache.{name}.slots"]))
    assert torch.equal(kc, t(gc[f"cache.{name}.kc_out"], dt))
    assert torch.equal(vc, t(gc[f"cache.{name}.vc_out"], dt))
    K.copy_blocks([kc], [vc], {0: [5, 6], 3: [7], 20: [1]})
    assert torch.equal(kc, t(gc[f"cache.{name}.kc_copied"], dt))
    assert torch.equal(vc, t(gc[f"cache.{name}.vc_copied"], dt))


@pytest.mark.parametrize("name", ["bfloat16", "float32"])
@pytest.mark.parametrize("heads", ["h4_4", "h8_2"])
@pytest.mark.parametrize("alibi", [0, 1])
def test_golden_cpu_attention(gc, name, heads, alibi):
    dt = DT[name]
    tag = f"attn.{name}.{heads}"
    q = t(gc[f"{tag}.q"], dt)
    ref = K.paged_attention(q, t(gc[f"{tag}.kc"], dt), t(gc[f"{tag}.vc"], dt),
                            torch.from_numpy(gc[f"{tag}.bt"]), torch.from_numpy(gc[f"{tag}.cl"]),
                            64 ** -0.5, t(gc[f"{tag}.alibi"], torch.float32) if alibi else None)
    for ver in ("v1", "v2"):
        got = t(gc[f"{tag}.{ver}.alibi{alibi}"], torch.float32)
        assert torch.allclose(got, ref, atol=1e-3, rtol=1e-5), (ver, (got - ref).abs().max())


# ---- 2. golden vectors from the reference's torch formulas --------------------------------------
@pytest.mark.parametrize("name", list(DT))
def test_golden_py_activation(gp, name):
    dt = DT[name]
    x = t(gp[f"act.{name}.x"], dt)
    for act in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
        out, ref = K.activation(x, act), t(gp[f"{act}.{name}"], dt)
        if act == "silu_and_mul" and dt != torch.float32:
            assert torch.equal(out, ref)           # tests/kernels/test_activation.py:49
        else:
            close_T(out, ref, dt)
    for act in ("gelu_new", "gelu_fast"):
        atol, rtol = TOL[dt]
        assert torch.allclose(K.activation(x, act).float(), t(gp[f"{act}.{name}"], torch.float32),
                              atol=max(atol, 8e-3 if dt != torch.float32 else 0), rtol=rtol)


@pytest.mark.parametrize("name", list(DT))
def test_golden_py_rms(gp, name):
    dt = DT[name]
    x, r, w = (t(gp[f"rms.{name}.{k}"], dt) for k in ("x", "r", "w"))
    assert torch.allclose(K.rms_norm(x, w, 1e-6), t(gp[f"rms.{name}.out"], dt), atol=1e-2, rtol=1e-2)
    ox, orr = K.fused_add_rms_norm(x, r, w, 1e-6)
    assert torch.allclose(ox, t(gp[f"frms.{name}.out_x"], dt), atol=1e-2, rtol=1e-2)
    assert torch.allclose(orr, t(gp[f"frms.{name}.out_r"], dt), atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize("name", list(DT))
@pytest.mark.parametrize("neox", [1, 0])
def test_golden_py_rope(gp, name, neox):
    dt = DT[name]
    tag = f"rope.{name}.neox{neox}"
    q, k = t(gp[f"{tag}.q"], dt), t(gp[f"{tag}.k"], dt)
    oq, ok = K.rotary_embedding(torch.from_numpy(gp[f"{tag}.pos"]), q.reshape(10, -1),
                                k.reshape(10, -1), 64, t(gp[f"{tag}.cache"], dt), bool(neox))
    atol, rtol = TOL[dt]
    assert torch.allclose(oq.reshape(q.shape), t(gp[f"{tag}.out_q"], dt), atol=atol, rtol=rtol)
    assert torch.allclose(ok.reshape(k.shape), t(gp[f"{tag}.out_k"], dt), atol=atol, rtol=rtol)


# ---- 3. live reference CPU kernels ---------------------------------------------------------------
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_live_ref_cpu(ref_cpu, dt):
    g = torch.Generator().manual_seed(77)
    x = torch.randn(33, 2 * 1024, generator=g).to(dt)
    o = torch.empty(33, 1024, dtype=dt)
    ref_cpu.ops.silu_and_mul(o, x)
    close_T(K.activation(x, "silu_and_mul"), o, dt)
    h = torch.randn(33, 4096, generator=g).to(dt)
    r = torch.randn(33, 4096, generator=g).to(dt)
    w = torch.randn(4096, generator=g).to(dt)
    ox, orr = K.fused_add_rms_norm(h, r, w, 1e-5)
    ref_cpu.ops.fused_add_rms_norm(h, r, w, 1e-5)
    assert torch.equal(orr, r)
    close_T(ox, h, dt)
    c = util.attention_case(5, (32, 8), 128, 16, dt, "auto", True, 3, 900, "cpu", num_blocks=256)
    out, _ = util.run_attention(ref_cpu.ops, c, "v1")
    assert torch.allclose(out.float(), util.oracle_attention(c), atol=1e-3, rtol=1e-5)
    out2, _ = util.run_attention(ref_cpu.ops, c, "v2")
    assert torch.allclose(out2.float(), util.oracle_attention(c), atol=1e-3, rtol=1e-5)


# ---- 4. fp8 e5m2 ---------------------------------------------------------------------------------
def test_fp8_e5m2_restatement_matches_torch_cast():
    allh = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.half)
    a, b = K.fp8_e5m2_from(allh), K.fp8_e5m2_from_fast(allh)
    nan = torch.isnan(allh)
    assert torch.equal(a[~nan], b[~nan])
    assert ((a[nan] & 0x7F) > 0x7C).all() and ((b[nan] & 0x7F) > 0x7C).all()
    # saturate-to-finite: inf and overflow map to +-57344 (0x7B / 0xFB), never to inf (0x7C)
    big = torch.tensor([1e9, -1e9, float("inf"), -float("inf"), 61439.0, 61440.0])
    assert K.fp8_e5m2_from(big).tolist() == [0x7B, 0xFB, 0x7B, 0xFB, 0x7B, 0x7B]
    g = torch.Generator().manual_seed(0)
    f = torch.randn(100000, generator=g) * torch.logspace(-9, 6, 100000)
    for dt in (torch.float32, torch.bfloat16):
        assert torch.equal(K.fp8_e5m2_from(f.to(dt)), K.fp8_e5m2_from_fast(f.to(dt)))


def test_fp8_e5m2_round_trip():
    codes = torch.arange(256, dtype=torch.uint8)
    h = K.fp8_e5m2_to(codes, torch.float16)
    fin = torch.isfinite(h)
    assert torch.equal(K.fp8_e5m2_from(h[fin]), codes[fin])   # every finite code is a fixed point
    assert torch.equal((h.view(torch.int16).to(torch.int32) & 0xFFFF)[fin],
                       (codes.to(torch.int32) << 8)[fin])      # e5m2 = top byte of half


def test_partition_oracle_is_consistent_with_full_oracle():
    c = util.attention_case(4, (8, 2), 64, 16, torch.float32, "auto", True, 5, 0, "cpu",
                            context_lens=[5, 512, 700, 1500], num_blocks=128)
    full = util.oracle_attention(c)
    ml, es, tmp = K.paged_attention_partitions(c["query"], c["key_cache"], c["value_cache"],
                                               c["block_tables"], c["context_lens"], c["scale"],
                                               512, c["alibi"])
    # merge the way paged_attention_v2_reduce_kernel does (attention_kernels.cu:511-602)
    live = ~torch.isnan(ml)
    mlz = torch.where(live, ml, torch.full_like(ml, -float("inf")))
    gm = mlz.max(dim=-1, keepdim=True).values
    wgt = torch.where(live, es * torch.exp(mlz - gm), torch.zeros_like(es))
    merged = (torch.nan_to_num(tmp) * (wgt / (wgt.sum(-1, keepdim=True) + 1e-6))[..., None]).sum(2)
    assert torch.allclose(merged, full, atol=1e-5, rtol=1e-4)


# ---- 5. prefill attention oracle vs torch SDPA (independent formulation) ------------------------
def test_context_attention_oracle_matches_sdpa():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    Hq, Hkv, D, bs = 8, 2, 64, 16
    sub, ctx = [5, 33, 1], [0, 40, 17]
    seq = [a + b for a, b in zip(sub, ctx)]
    T = sum(sub)
    q = torch.randn(T, Hq, D, generator=g)
    k, v = torch.randn(T, Hkv, D, generator=g), torch.randn(T, Hkv, D, generator=g)
    full_k = [torch.randn(s, Hkv, D, generator=g) for s in seq]
    full_v = [torch.randn(s, Hkv, D, generator=g) for s in seq]
    mb = max((s + bs - 1) // bs for s in seq)
    nb = 3 * mb
    kc = torch.zeros(nb, Hkv, D // 4, bs, 4)
    vc = torch.zeros(nb, Hkv, D, bs)
    bt = torch.arange(nb, dtype=torch.int32).view(3, mb)
    start = torch.tensor([0, 5, 38], dtype=torch.int32)
    for b in range(3):
        s0 = int(start[b])
        full_k[b][ctx[b]:] = k[s0:s0 + sub[b]]
        full_v[b][ctx[b]:] = v[s0:s0 + sub[b]]
        for pos in range(ctx[b]):
            blk, off = int(bt[b, pos // bs]), pos % bs
            kc[blk, :, :, off, :] = full_k[b][pos].view(Hkv, D // 4, 4)
            vc[blk, :, :, off] = full_v[b][pos]
    out = K.context_attention(q, k, v, kc, vc, bt, start, torch.tensor(seq), torch.tensor(ctx))
    for b in range(3):
        s0, n, c = int(start[b]), sub[b], ctx[b]
        qq = q[s0:s0 + n].transpose(0, 1)                                   # [Hq, n, D]
        kk = full_k[b].repeat_interleave(Hq // Hkv, dim=1).transpose(0, 1)  # [Hq, S, D]
        vv = full_v[b].repeat_interleave(Hq // Hkv, dim=1).transpose(0, 1)
        mask = torch.arange(c + n)[None, :] <= (c + torch.arange(n))[:, None]
        ref = F.scaled_dot_product_attention(qq, kk, vv, attn_mask=mask).transpose(0, 1)
        assert torch.allclose(out[s0:s0 + n], ref, atol=1e-5, rtol=1e-4)
