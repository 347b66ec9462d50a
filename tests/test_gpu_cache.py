"""GPU parity of the KV-cache index/copy ops — bit-exact (torch.equal) against the oracle.
Grids follow tests/kernels/test_cache.py:30-262 of the reference; the reference skips the fp8
rows of reshape_and_cache / swap_blocks on CUDA, here they run and are exact too (the oracle
restates the e5m2 saturate-to-finite conversion)."""
import random

import pytest
import torch

from tests import util
from oracle import kernels as K

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
DTYPES = [torch.half, torch.bfloat16, torch.float]
HEAD_SIZES = [64, 80, 96, 112, 128, 256]
BLOCK_SIZES = [8, 16, 32]


@pytest.fixture(scope="module")
def cache_ops():
    from vllm_ltr_b200 import _C
    return _C.cache_ops


def _caches(num_blocks, block_size, layers, heads, head_size, dtype, kv, seed, device=DEV):
    g = torch.Generator(device=device).manual_seed(seed)
    ks, vs = [], []
    for _ in range(layers):
        k, v = util.make_cache(num_blocks, block_size, heads, head_size, dtype, kv, g, device)
        ks.append(k)
        vs.append(v)
    return ks, vs


@pytest.mark.parametrize("head_size", HEAD_SIZES)
@pytest.mark.parametrize("block_size", BLOCK_SIZES)
@pytest.mark.parametrize("num_blocks", [1024, 10000])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kv", ["auto", "fp8"])
@pytest.mark.parametrize("num_layers", [1, 5])
def test_copy_blocks(cache_ops, head_size, block_size, num_blocks, dtype, kv, num_layers):
    if num_layers > 1 and (num_blocks > 1024 or head_size not in (64, 128)):
        pytest.skip("multi-layer rows only on a sub-grid")
    rnd = random.Random(0)
    n = 256
    src = rnd.sample(range(num_blocks), n)
    rest = list(set(range(num_blocks)) - set(src))
    dst = rnd.sample(rest, 2 * n)
    mapping = {src[i]: [dst[2 * i], dst[2 * i + 1]] for i in range(n)}
    ks, vs = _caches(num_blocks, block_size, num_layers, 8, head_size, dtype, kv, 0)
    rk = [k.cpu() for k in ks]
    rv = [v.cpu() for v in vs]
    cache_ops.copy_blocks(ks, vs, mapping)
    K.copy_blocks(rk, rv, mapping)
    for a, b in zip(ks + vs, rk + rv):
        assert torch.equal(a.cpu(), b)


def test_copy_blocks_many_layers_and_pairs(cache_ops):
    """80 layers (70B) x 700 pairs: more than one launch's worth of kernel-parameter space."""
    rnd = random.Random(1)
    nb, L = 2048, 80
    src = rnd.sample(range(nb), 700)
    rest = list(set(range(nb)) - set(src))
    dst = rnd.sample(rest, 700)
    mapping = {s: [d] for s, d in zip(src, dst)}
    ks, vs = _caches(nb, 16, L, 1, 128, torch.bfloat16, "auto", 2)
    rk = [k.cpu() for k in ks]
    rv = [v.cpu() for v in vs]
    cache_ops.copy_blocks(ks, vs, mapping)
    K.copy_blocks(rk, rv, mapping)
    for a, b in zip(ks + vs, rk + rv):
        assert torch.equal(a.cpu(), b)
    cache_ops.copy_blocks(ks, vs, {})       # empty mapping is a no-op
    cache_ops.copy_blocks([], [], {1: [2]})


@pytest.mark.parametrize("head_size", HEAD_SIZES)
@pytest.mark.parametrize("block_size", BLOCK_SIZES)
@pytest.mark.parametrize("num_blocks", [1024, 10000])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kv", ["auto", "fp8"])
def test_reshape_and_cache(cache_ops, head_size, block_size, num_blocks, dtype, kv):
    rnd = random.Random(0)
    T, H = 42, 8
    slots = rnd.sample(range(num_blocks * block_size), T)
    slots[5] = -1                                   # padding token: skipped (cache_kernels.cu:172)
    slot_mapping = torch.tensor(slots, dtype=torch.long, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    qkv = torch.randn(T, 3, H, head_size, generator=g, device=DEV).to(dtype)
    if kv == "fp8":
        qkv[0, 1, 0, :4] = torch.tensor([1e6, -1e6, 60000.0, 3e-6], device=DEV).to(dtype)
    _, key, value = qkv.unbind(dim=1)               # strided views, like the fused qkv output
    ks, vs = _caches(num_blocks, block_size, 1, H, head_size, dtype, kv, 0)
    kc, vc = ks[0], vs[0]
    rk, rv = kc.cpu(), vc.cpu()
    cache_ops.reshape_and_cache(key, value, kc, vc, slot_mapping, kv, 1.0)
    K.reshape_and_cache(key.cpu(), value.cpu(), rk, rv, slot_mapping.cpu(), kv)
    assert torch.equal(kc.cpu(), rk)
    assert torch.equal(vc.cpu(), rv)


@pytest.mark.parametrize("num_tokens", [1, 256, 768])
@pytest.mark.parametrize("heads,dtype,kv", [(32, torch.half, "auto"), (8, torch.bfloat16, "auto"),
                                             (8, torch.bfloat16, "fp8"), (1, torch.bfloat16, "auto")])
def test_reshape_and_cache_engine_shapes(cache_ops, num_tokens, heads, dtype, kv):
    """Engine shapes: key/value are column slices of the fused qkv GEMM output."""
    rnd = random.Random(4)
    D, bs, nb, hq = 128, 16, 4096, 32
    g = torch.Generator(device=DEV).manual_seed(4)
    qkv = torch.randn(num_tokens, (hq + 2 * heads) * D, generator=g, device=DEV).to(dtype)
    key = qkv[:, hq * D:(hq + heads) * D].view(num_tokens, heads, D)
    value = qkv[:, (hq + heads) * D:].view(num_tokens, heads, D)
    slots = torch.tensor(rnd.sample(range(nb * bs), num_tokens), dtype=torch.long, device=DEV)
    ks, vs = _caches(nb, bs, 1, heads, D, dtype, kv, 4)
    rk, rv = ks[0].cpu(), vs[0].cpu()
    cache_ops.reshape_and_cache(key, value, ks[0], vs[0], slots, kv, 1.0)
    K.reshape_and_cache(key.cpu(), value.cpu(), rk, rv, slots.cpu(), kv)
    assert torch.equal(ks[0].cpu(), rk) and torch.equal(vs[0].cpu(), rv)


def _run_slots(rnd, bs, nb):
    """Slot mapping of a mixed engine step: prefill chunks (runs of consecutive slots that start
    and end anywhere inside a block and continue in unrelated blocks), padding inside a run,
    a descending pair, a run broken by one slot, then lone decode tokens."""
    blocks = rnd.sample(range(nb), 64)
    slots = []
    for start, length in ((0, 3 * bs), (bs - 3, bs + 5), (5, 2), (bs - 1, 1), (1, 4 * bs - 2)):
        pos = start
        for _ in range(length):
            slots.append(blocks[pos // bs] * bs + pos % bs)
            pos += 1
        blocks = blocks[8:]
    b = blocks[0] * bs
    slots += [b + 2, b + 3, -1, b + 5, b + 6]            # padding token inside a run
    b = blocks[1] * bs
    slots += [b + 7, b + 6, b + 5]                       # descending: three runs of one
    b = blocks[2] * bs
    slots += [b, b + 1, b + 3, b + 4, b + 4 + bs]        # a gap, then the same offset one block on
    slots += [blk * bs + rnd.randrange(bs) for blk in blocks[3:20]]
    return slots


@pytest.mark.parametrize("block_size", BLOCK_SIZES)
@pytest.mark.parametrize("head_size", [64, 80, 128])
@pytest.mark.parametrize("heads,dtype,kv", [(32, torch.half, "auto"), (8, torch.bfloat16, "auto"),
                                             (8, torch.bfloat16, "fp8"), (8, torch.half, "fp8"),
                                             (2, torch.float, "auto"), (2, torch.float, "fp8")])
def test_reshape_and_cache_runs(cache_ops, block_size, head_size, heads, dtype, kv):
    """Tokens with consecutive slots inside a block are written by one warp with wide stores
    (csrc/cache.cu): every way a run can start, end and break must give the reference's bytes."""
    rnd = random.Random(7)
    nb, hq = 512, 4
    slots_l = _run_slots(rnd, block_size, nb)
    T = len(slots_l)
    g = torch.Generator(device=DEV).manual_seed(7)
    qkv = torch.randn(T, (hq + 2 * heads) * head_size, generator=g, device=DEV).to(dtype)
    key = qkv[:, hq * head_size:(hq + heads) * head_size].view(T, heads, head_size)
    value = qkv[:, (hq + heads) * head_size:].view(T, heads, head_size)
    slots = torch.tensor(slots_l, dtype=torch.long, device=DEV)
    ks, vs = _caches(nb, block_size, 1, heads, head_size, dtype, kv, 7)
    rk, rv = ks[0].cpu(), vs[0].cpu()
    cache_ops.reshape_and_cache(key, value, ks[0], vs[0], slots, kv, 1.0)
    K.reshape_and_cache(key.cpu(), value.cpu(), rk, rv, slots.cpu(), kv)
    assert torch.equal(ks[0].cpu(), rk) and torch.equal(vs[0].cpu(), rv)


def test_reshape_and_cache_prefill_step(cache_ops):
    """A 768-token engine step: two prompt chunks and 200 decode tokens, Llama-2-7B heads."""
    rnd = random.Random(9)
    D, bs, nb, H = 128, 16, 4096, 32
    blocks = rnd.sample(range(nb), 300)
    slots_l = [blocks[i // bs] * bs + i % bs for i in range(400)]                  # new prompt
    slots_l += [blocks[40 + (9 + i) // bs] * bs + (9 + i) % bs for i in range(168)]  # continued
    slots_l += [blocks[100 + i] * bs + rnd.randrange(bs) for i in range(200)]      # decode
    T = len(slots_l)
    g = torch.Generator(device=DEV).manual_seed(9)
    qkv = torch.randn(T, 3 * H * D, generator=g, device=DEV).to(torch.half)
    key = qkv[:, H * D:2 * H * D].view(T, H, D)
    value = qkv[:, 2 * H * D:].view(T, H, D)
    slots = torch.tensor(slots_l, dtype=torch.long, device=DEV)
    ks, vs = _caches(nb, bs, 1, H, D, torch.half, "auto", 9)
    rk, rv = ks[0].cpu(), vs[0].cpu()
    cache_ops.reshape_and_cache(key, value, ks[0], vs[0], slots, "auto", 1.0)
    K.reshape_and_cache(key.cpu(), value.cpu(), rk, rv, slots.cpu(), "auto")
    assert torch.equal(ks[0].cpu(), rk) and torch.equal(vs[0].cpu(), rv)


@pytest.mark.parametrize("direction", [("cuda", "cpu"), ("cuda", "cuda"), ("cpu", "cuda")])
@pytest.mark.parametrize("head_size", [64, 80, 128, 256])
@pytest.mark.parametrize("block_size", BLOCK_SIZES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kv", ["auto", "fp8"])
@pytest.mark.parametrize("pinned", [True, False])
def test_swap_blocks(cache_ops, direction, head_size, block_size, dtype, kv, pinned):
    """Pinned host memory (what CacheEngine allocates, worker/cache_engine.py:70-77) goes through
    the single gather/scatter kernel; pageable host memory through coalesced memcpys."""
    if "cpu" not in direction and not pinned:
        pytest.skip("pinning only matters for host memory")
    rnd = random.Random(0)
    nb, n = 1024, 256
    src_dev = DEV if direction[0] == "cuda" else "cpu"
    dst_dev = DEV if direction[1] == "cuda" else "cpu"
    src_blocks = rnd.sample(range(nb), n)
    if src_dev == dst_dev:
        dst_blocks = rnd.sample(list(set(range(nb)) - set(src_blocks)), n)
    else:
        dst_blocks = rnd.sample(range(nb), n)
    mapping = dict(zip(src_blocks, dst_blocks))
    sk, sv = _caches(nb, block_size, 1, 8, head_size, dtype, kv, 0, src_dev)
    dk, dv = _caches(nb, block_size, 1, 8, head_size, dtype, kv, 1, dst_dev)
    if pinned:
        if src_dev == "cpu":
            sk, sv = [sk[0].pin_memory()], [sv[0].pin_memory()]
        if dst_dev == "cpu":
            dk, dv = [dk[0].pin_memory()], [dv[0].pin_memory()]
    if src_dev == dst_dev:
        dk, dv = sk, sv        # the reference swaps within one cache tensor when on one device
    ek, ev = dk[0].cpu().clone(), dv[0].cpu().clone()
    K.swap_blocks(sk[0].cpu(), ek, mapping)
    K.swap_blocks(sv[0].cpu(), ev, mapping)
    cache_ops.swap_blocks(sk[0], dk[0], mapping)
    cache_ops.swap_blocks(sv[0], dv[0], mapping)
    torch.cuda.synchronize()
    assert torch.equal(dk[0].cpu(), ek)
    assert torch.equal(dv[0].cpu(), ev)


@pytest.mark.parametrize("direction", [("cuda", "cpu"), ("cpu", "cuda"), ("cuda", "cuda")])
@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("kv", ["auto", "fp8"])
def test_swap_blocks_layers(cache_ops, direction, pinned, kv):
    """All layers of CacheEngine.swap_in / swap_out in one call (cache_engine.py:92-100) against
    the reference loop: per layer, swap_blocks on the K half and on the V half
    (paged_attn.py:195-206)."""
    if "cpu" not in direction and not pinned:
        pytest.skip("pinning only matters for host memory")
    rnd = random.Random(1)
    layers, nb, n, bs, hkv, D = 5, 128, 48, 16, 4, 64
    dtype = torch.uint8 if kv == "fp8" else torch.half
    g = torch.Generator().manual_seed(3)

    def make(dev, seed_off):
        out = []
        for _ in range(layers):
            t = torch.randint(0, 255, (2, nb, bs * hkv * D), generator=g, dtype=torch.uint8)
            t = t if dtype == torch.uint8 else t.to(torch.int16).view(torch.half).clone()
            t = t.to(DEV) if dev == "cuda" else (t.pin_memory() if pinned else t)
            out.append(t)
        return out

    src, dst = make(direction[0], 0), make(direction[1], 1)
    src_blocks = rnd.sample(range(nb), n)
    dst_blocks = rnd.sample(range(nb), n)
    if direction == ("cuda", "cuda"):
        dst = src
        dst_blocks = rnd.sample(list(set(range(nb)) - set(src_blocks)), n)
    mapping = dict(zip(src_blocks, dst_blocks))
    want = [d.cpu().clone() for d in dst]
    for s, w in zip(src, want):
        K.swap_blocks(s[0].cpu(), w[0], mapping)
        K.swap_blocks(s[1].cpu(), w[1], mapping)
    cache_ops.swap_blocks_layers(src, dst, mapping)
    torch.cuda.synchronize()
    for d, w in zip(dst, want):
        assert torch.equal(d.cpu().view(torch.uint8), w.view(torch.uint8))


def test_swap_blocks_errors(cache_ops):
    a = torch.zeros(4, 8, dtype=torch.half)
    with pytest.raises(RuntimeError, match="Invalid device combination"):
        cache_ops.swap_blocks(a, a.clone(), {0: 1})


@pytest.mark.parametrize("dtype", DTYPES)
def test_convert_fp8(cache_ops, dtype):
    """to-fp8 over every half bit pattern (incl. inf/NaN/subnormals/overflow) and back."""
    allh = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.half)
    if dtype == torch.half:
        src = allh.clone()
    else:
        g = torch.Generator().manual_seed(0)
        extra = torch.randn(65536, generator=g) * torch.logspace(-8, 6, 65536)
        src = torch.cat([allh.float(), extra]).to(dtype)
    dst = torch.empty(src.shape, dtype=torch.uint8, device=DEV)
    cache_ops.convert_fp8(src.to(DEV), dst)
    want = K.fp8_e5m2_from(src)
    got = dst.cpu()
    nan = torch.isnan(src.float())
    assert torch.equal(got[~nan], want[~nan])
    assert ((got[nan] & 0x7F) > 0x7C).all()         # NaN stays NaN
    back = torch.empty(256, dtype=dtype, device=DEV)
    codes = torch.arange(256, dtype=torch.uint8)
    cache_ops.convert_fp8(codes.to(DEV), back)
    ref = K.fp8_e5m2_to(codes, dtype)
    fin = ~torch.isnan(ref.float())
    assert torch.equal(back.cpu()[fin], ref[fin])
    assert torch.isnan(back.cpu()[~fin].float()).all()


def test_cache_ops_match_reference_cuda_build(cache_ops, ref_cuda):
    """reshape_and_cache / copy_blocks / convert_fp8 bit-exact against the reference's own CUDA
    kernels recompiled for sm_100a."""
    rnd = random.Random(9)
    for dtype in DTYPES:
        for kv in ("auto", "fp8"):
            T, H, D, bs, nb = 300, 8, 128, 16, 512
            g = torch.Generator(device=DEV).manual_seed(9)
            qkv = (torch.randn(T, 3, H, D, generator=g, device=DEV) * 4).to(dtype)
            _, key, value = qkv.unbind(dim=1)
            slots = torch.tensor(rnd.sample(range(nb * bs), T), dtype=torch.long, device=DEV)
            ks, vs = _caches(nb, bs, 2, H, D, dtype, kv, 9)
            k2 = [k.clone() for k in ks]
            v2 = [v.clone() for v in vs]
            cache_ops.reshape_and_cache(key, value, ks[0], vs[0], slots, kv, 1.0)
            ref_cuda.cache_ops.reshape_and_cache(key, value, k2[0], v2[0], slots, kv, 1.0)
            mapping = {1: [2, 3], 7: [100], 500: [0]}
            cache_ops.copy_blocks(ks, vs, mapping)
            ref_cuda.cache_ops.copy_blocks(k2, v2, mapping)
            for a, b in zip(ks + vs, k2 + v2):
                assert torch.equal(a, b)
