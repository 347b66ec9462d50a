"""GPU: the fused rotary-embedding + cache-append launch (csrc/rope_cache.cu) against the two ops
it replaces, through the C ABI: ``ops.rotary_embedding`` followed by ``cache_ops.reshape_and_cache``
(each of which has its own parity tests against the oracle and the reference's CUDA kernels).
Everything is index / copy work plus the reference's "in T" rotation, so the comparison is
bit-exact: rotated queries, rotated keys (written back in place) and both caches."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cos_sin(head_size, max_pos, dtype, base=10000.0):
    inv = 1.0 / (base ** (torch.arange(0, head_size, 2, dtype=torch.float) / head_size))
    f = torch.einsum("i,j->ij", torch.arange(max_pos, dtype=torch.float), inv)
    return torch.cat((f.cos(), f.sin()), dim=-1).to(dtype).to(DEV)


def _slots(rnd, bs, nb):
    """prefill runs that start / end anywhere in a block, padding, lone decode tokens"""
    blocks = rnd.sample(range(nb), 40)
    out = []
    for start, length in ((0, 2 * bs + 3), (bs - 2, bs + 1), (3, 1)):
        for p in range(start, start + length):
            out.append(blocks[p // bs] * bs + p % bs)
        blocks = blocks[6:]
    b = blocks[0] * bs
    out += [b + 1, b + 2, -1, b + 4, -1]
    out += [blk * bs + rnd.randrange(bs) for blk in blocks[1:12]]
    return out


@pytest.mark.parametrize("block_size", [8, 16, 32])
@pytest.mark.parametrize("head_size", [64, 80, 128, 256])
@pytest.mark.parametrize("heads,dtype,kv", [((32, 32), torch.half, "auto"),
                                             ((32, 8), torch.bfloat16, "auto"),
                                             ((32, 8), torch.bfloat16, "fp8"),
                                             ((8, 1), torch.half, "fp8"),
                                             ((4, 2), torch.float, "auto"),
                                             ((4, 2), torch.float, "fp8")])
def test_fused_matches_the_two_ops(block_size, head_size, heads, dtype, kv):
    from vllm_ltr_b200 import _C
    from vllm_ltr_b200.attention_backend import rotary_embedding_and_cache
    rnd = random.Random(block_size * 1000 + head_size)
    hq, hkv = heads
    nb, max_pos = 256, 4096
    slots_l = _slots(rnd, block_size, nb)
    T = len(slots_l)
    g = torch.Generator(device=DEV).manual_seed(T + head_size)
    qkv = torch.randn(T, (hq + 2 * hkv) * head_size, generator=g, device=DEV).to(dtype)
    q, k, v = qkv.split([hq * head_size, hkv * head_size, hkv * head_size], dim=-1)
    pos = torch.randint(0, max_pos, (T,), generator=g, device=DEV)
    slots = torch.tensor(slots_l, dtype=torch.long, device=DEV)
    cs = _cos_sin(head_size, max_pos, dtype)
    es = 1 if kv == "fp8" else qkv.element_size()
    x = 16 // es
    cdt = torch.uint8 if kv == "fp8" else dtype
    kc = torch.zeros(nb, hkv, head_size // x, block_size, x, dtype=cdt, device=DEV)
    vc = torch.zeros(nb, hkv, head_size, block_size, dtype=cdt, device=DEV)
    # the reference sequence on a copy
    qkv2 = qkv.clone()
    q2, k2, v2 = qkv2.split([hq * head_size, hkv * head_size, hkv * head_size], dim=-1)
    kc2, vc2 = kc.clone(), vc.clone()
    _C.ops.rotary_embedding(pos, q2, k2, head_size, cs, True)
    _C.cache_ops.reshape_and_cache(k2.view(T, hkv, head_size), v2.view(T, hkv, head_size), kc2, vc2,
                                   slots, kv, 1.0)
    fused = rotary_embedding_and_cache(pos, q, k, v, kc, vc, slots, head_size, cs, True, kv)
    if kv == "fp8" and block_size == 8:
        # a V row of 8 e5m2 tokens is half a 16-byte group: outside the fused form (the separate
        # op takes its scalar path there too)
        assert not fused
        return
    assert fused
    assert torch.equal(qkv.view(torch.uint8), qkv2.view(torch.uint8))     # q, k rotated in place; v untouched
    assert torch.equal(kc.view(torch.uint8), kc2.view(torch.uint8))
    assert torch.equal(vc.view(torch.uint8), vc2.view(torch.uint8))


def test_engine_step_shapes():
    """A 768-token mixed step of Llama-2-7B geometry (512-token chunk + 256 decode tokens)."""
    from vllm_ltr_b200 import _C
    from vllm_ltr_b200.attention_backend import rotary_embedding_and_cache
    rnd = random.Random(3)
    H, D, bs, nb = 32, 128, 16, 4096
    blocks = rnd.sample(range(nb), 400)
    sl = [blocks[(7 + i) // bs] * bs + (7 + i) % bs for i in range(512)]
    sl += [blocks[100 + i] * bs + rnd.randrange(bs) for i in range(256)]
    T = len(sl)
    g = torch.Generator(device=DEV).manual_seed(3)
    qkv = torch.randn(T, 3 * H * D, generator=g, device=DEV).half()
    q, k, v = qkv.split([H * D] * 3, dim=-1)
    pos = torch.randint(0, 4096, (T,), generator=g, device=DEV)
    slots = torch.tensor(sl, dtype=torch.long, device=DEV)
    cs = _cos_sin(D, 4096, torch.half)
    kc = torch.zeros(nb, H, D // 8, bs, 8, dtype=torch.half, device=DEV)
    vc = torch.zeros(nb, H, D, bs, dtype=torch.half, device=DEV)
    qkv2, kc2, vc2 = qkv.clone(), kc.clone(), vc.clone()
    q2, k2, v2 = qkv2.split([H * D] * 3, dim=-1)
    _C.ops.rotary_embedding(pos, q2, k2, D, cs, True)
    _C.cache_ops.reshape_and_cache(k2.view(T, H, D), v2.view(T, H, D), kc2, vc2, slots, "auto", 1.0)
    assert rotary_embedding_and_cache(pos, q, k, v, kc, vc, slots, D, cs, True, "auto")
    assert torch.equal(qkv, qkv2) and torch.equal(kc, kc2) and torch.equal(vc, vc2)


def test_outside_the_fused_form_nothing_is_touched():
    from vllm_ltr_b200.attention_backend import rotary_embedding_and_cache
    H, D, bs, nb, T = 4, 128, 16, 8, 5
    qkv = torch.randn(T, 3 * H * D, device=DEV).half()
    q, k, v = qkv.split([H * D] * 3, dim=-1)
    pos = torch.arange(T, device=DEV)
    slots = torch.arange(T, device=DEV)
    kc = torch.zeros(nb, H, D // 8, bs, 8, dtype=torch.half, device=DEV)
    vc = torch.zeros(nb, H, D, bs, dtype=torch.half, device=DEV)
    before = qkv.clone()
    # GPT-J style rotation and partial rotation are not fused
    assert not rotary_embedding_and_cache(pos, q, k, v, kc, vc, slots, D, _cos_sin(D, 64, torch.half),
                                          False, "auto")
    assert not rotary_embedding_and_cache(pos, q, k, v, kc, vc, slots, D, _cos_sin(64, 64, torch.half),
                                          True, "auto")
    # mismatched cos/sin dtype
    assert not rotary_embedding_and_cache(pos, q, k, v, kc, vc, slots, D,
                                          _cos_sin(D, 64, torch.float), True, "auto")
    assert torch.equal(qkv, before) and not kc.any() and not vc.any()
    with pytest.raises(RuntimeError, match="no CPU path"):
        rotary_embedding_and_cache(pos.cpu(), q.cpu(), k.cpu(), v.cpu(), kc.cpu(), vc.cpu(),
                                   slots.cpu(), D, _cos_sin(D, 64, torch.half).cpu(), True)
    # empty step
    assert rotary_embedding_and_cache(pos[:0], q[:0], k[:0], v[:0], kc, vc, slots[:0], D,
                                      _cos_sin(D, 64, torch.half), True)
