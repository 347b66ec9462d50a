"""Size-independent properties of the hot path at the benchmark's FULL sizes (B = 256 running
sequences of the synthetic ShareGPT trace, Llama-2-7B / Llama-3-8B head shapes, block 16), where
the float64 oracle would take minutes: relocation invariance, exact power-of-two linearity,
v1 == v2, cache round trips, swap-out/swap-in round trips, copy idempotence, rank sortedness."""
import numpy as np
import pytest
import torch

from bench import build_inputs, trace_contexts

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def C():
    from vllm_ltr_b200 import _C
    return _C


def _case(hq, hkv, dtype, kv, B=256, seed=0):
    D, bs = 128, 16
    ctx = trace_contexts(B, seed)
    need = int(((ctx + bs - 1) // bs).sum())
    nb = need + 64
    inp = build_inputs(ctx, bs, nb, seed)
    g = torch.Generator(device=DEV).manual_seed(seed)
    es = 1 if kv == "fp8" else 2
    x = 16 // es
    if kv == "fp8":
        kc = torch.randint(0, 0x38, (nb, hkv, D // x, bs, x), generator=g, device=DEV, dtype=torch.uint8)
        vc = torch.randint(0, 0x38, (nb, hkv, D, bs), generator=g, device=DEV, dtype=torch.uint8)
    else:
        kc = ((torch.rand(nb, hkv, D // x, bs, x, generator=g, device=DEV) * 2 - 1) * D ** -0.5).to(dtype)
        vc = ((torch.rand(nb, hkv, D, bs, generator=g, device=DEV) * 2 - 1) * D ** -0.5).to(dtype)
    q = ((torch.rand(B, hq, D, generator=g, device=DEV) * 2 - 1) * D ** -0.5).to(dtype)
    return dict(q=q, kc=kc, vc=vc, bt=inp["block_tables"].to(DEV), cl=inp["context_lens"].to(DEV),
                mx=int(ctx.max()), hkv=hkv, kv=kv, nb=nb, need=need)


def _attn(C, c, version="v1", kc=None, vc=None, bt=None):
    q = c["q"]
    kc = c["kc"] if kc is None else kc
    vc = c["vc"] if vc is None else vc
    bt = c["bt"] if bt is None else bt
    out = torch.empty_like(q)
    if version == "v1":
        C.ops.paged_attention_v1(out, q, kc, vc, c["hkv"], 128 ** -0.5, bt, c["cl"], 16, c["mx"],
                                 None, c["kv"], 1.0)
    else:
        B, H, D = q.shape
        P = (c["mx"] + 511) // 512
        tmp = torch.empty(B, H, P, D, dtype=q.dtype, device=DEV)
        es = torch.empty(B, H, P, dtype=torch.float32, device=DEV)
        C.ops.paged_attention_v2(out, es, torch.empty_like(es), tmp, q, kc, vc, c["hkv"],
                                 128 ** -0.5, bt, c["cl"], 16, c["mx"], None, c["kv"], 1.0)
    return out


CONFIGS = [(32, 32, torch.float16, "auto"), (32, 8, torch.bfloat16, "auto"),
           (32, 8, torch.bfloat16, "fp8"), (8, 1, torch.bfloat16, "auto")]


@pytest.mark.parametrize("hq,hkv,dtype,kv", CONFIGS)
def test_attention_properties_full_size(C, hq, hkv, dtype, kv):
    c = _case(hq, hkv, dtype, kv)
    base = _attn(C, c)
    assert torch.isfinite(base.float()).all()
    # (1) deterministic: the work-queue order must not leak into the result
    assert torch.equal(base, _attn(C, c))
    # (2) relocation invariance: move every used block to a new physical slot, remap the table
    perm = torch.randperm(c["nb"], device=DEV)
    kc2, vc2 = torch.empty_like(c["kc"]), torch.empty_like(c["vc"])
    kc2[perm] = c["kc"]
    vc2[perm] = c["vc"]
    bt2 = perm[c["bt"].long()].to(torch.int32)
    assert torch.equal(base, _attn(C, c, kc=kc2, vc=vc2, bt=bt2))
    # (3) v1 and v2 partition the contexts differently: same result within the fp tolerance
    tol = 1e-2 if kv == "fp8" else 1e-3
    assert torch.allclose(base.float(), _attn(C, c, "v2").float(), atol=tol, rtol=1e-5)
    # (4) exact linearity in V under power-of-two scaling (2-byte caches: exponent shift)
    if kv == "auto":
        doubled = _attn(C, c, vc=c["vc"] * 2).float()
        normal = base.float().abs() >= 2.0 ** -13      # f16 subnormal results round on another grid
        assert torch.equal(doubled[normal], (base.float() * 2)[normal])
        assert torch.allclose(doubled, base.float() * 2, atol=2.0 ** -23, rtol=0)
    # (5) a convex combination of V rows: every output lies inside the value range
    lo, hi = -128 ** -0.5, 128 ** -0.5
    if kv == "auto":
        assert base.float().min() >= lo - 1e-3 and base.float().max() <= hi + 1e-3


@pytest.mark.parametrize("hkv,dtype,kv", [(32, torch.float16, "auto"), (8, torch.bfloat16, "auto"),
                                           (8, torch.bfloat16, "fp8")])
def test_cache_round_trips_full_size(C, hkv, dtype, kv):
    """reshape_and_cache -> read the slots back == the keys/values written (after e5m2 rounding
    for fp8); copy_blocks is idempotent; swap out + swap in restores the blocks bit for bit."""
    c = _case(32, hkv, dtype, kv, seed=3)
    B, D, bs = 256, 128, 16
    g = torch.Generator(device=DEV).manual_seed(1)
    key = torch.randn(B, hkv, D, generator=g, device=DEV).to(dtype)
    val = torch.randn(B, hkv, D, generator=g, device=DEV).to(dtype)
    ctx = c["cl"].long()
    pos = ctx - 1
    slots = c["bt"].long()[torch.arange(B, device=DEV), pos // bs] * bs + pos % bs
    kc, vc = c["kc"].clone(), c["vc"].clone()
    C.cache_ops.reshape_and_cache(key, val, kc, vc, slots, kv, 1.0)
    blk, off = slots // bs, slots % bs
    k_back = kc[blk, :, :, off, :].reshape(B, hkv, D)
    v_back = vc[blk, :, :, off]
    if kv == "fp8":
        from oracle import kernels as K
        assert torch.equal(k_back, K.fp8_e5m2_from_fast(key)) and torch.equal(v_back, K.fp8_e5m2_from_fast(val))
    else:
        assert torch.equal(k_back, key) and torch.equal(v_back, val)
    # untouched slots keep their bytes
    mask = torch.ones(c["nb"] * bs, dtype=torch.bool, device=DEV)
    mask[slots] = False
    assert torch.equal(vc.permute(0, 3, 1, 2).reshape(-1, hkv, D)[mask],
                       c["vc"].permute(0, 3, 1, 2).reshape(-1, hkv, D)[mask])
    # copy_blocks twice == once
    mapping = {int(i): [int(c["nb"] - 1 - i)] for i in range(0, 32)}
    k1, v1 = kc.clone(), vc.clone()
    C.cache_ops.copy_blocks([k1], [v1], mapping)
    k2, v2 = k1.clone(), v1.clone()
    C.cache_ops.copy_blocks([k2], [v2], mapping)
    assert torch.equal(k1, k2) and torch.equal(v1, v2)
    # swap out to pinned host memory and back in to different GPU blocks
    host = torch.empty_like(kc[:64], device="cpu").pin_memory()
    out_map = {int(b): i for i, b in enumerate(range(100, 164))}
    C.cache_ops.swap_blocks(kc, host, out_map)
    in_map = {i: int(b) for i, b in enumerate(range(300, 364))}
    kc3 = kc.clone()
    C.cache_ops.swap_blocks(host, kc3, in_map)
    torch.cuda.synchronize()
    assert torch.equal(kc3[300:364], kc[100:164])


def test_rank_properties_full_size():
    from vllm_ltr_b200.ltr_scheduler import GpuRanker
    r = GpuRanker()
    rng = np.random.default_rng(1)
    s = rng.normal(0, 1, 10000).astype(np.float16).astype(np.float64)
    pri = rng.choice([-1, 0], 10000, p=[0.05, 0.95]).astype(np.int32)
    o = r.order(s, pri)
    assert np.array_equal(np.sort(o), np.arange(10000))
    key = np.stack([pri[o].astype(np.float64), -s[o]], 1)
    assert all(tuple(a) <= tuple(b) for a, b in zip(key[:-1], key[1:]))
    assert np.array_equal(r.order(s, pri), o)                      # idempotent
    o2 = r.order(s[o], pri[o])                                      # already ordered: identity
    assert np.array_equal(o2, np.arange(10000))
