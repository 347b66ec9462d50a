"""GPU parity of rms_norm / fused_add_rms_norm / rotary_embedding / activations through the
vllm._C mirror.  Grids and tolerances are those of the reference's tests/kernels/
{test_layernorm,test_pos_encoding,test_activation}.py; on top of them f16/bf16 element-wise ops are
checked BIT-EXACTLY against the oracle and against the reference's CUDA kernels (oracle/_ref)."""
import pytest
import torch
import torch.nn.functional as F

from tests import util
from oracle import kernels as K

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
DTYPES = [torch.half, torch.bfloat16, torch.float]
TOL = {torch.half: (1e-3, 1e-3), torch.bfloat16: (1e-3, 1.6e-2), torch.float: (1e-5, 1.3e-6)}


@pytest.fixture(scope="module")
def ops():
    from vllm_ltr_b200 import _C
    return _C.ops


# ---- layernorm (tests/kernels/test_layernorm.py:17-54) -------------------------------------------
@pytest.mark.parametrize("num_tokens", [1, 7, 83, 4096])
@pytest.mark.parametrize("hidden_size", [8, 768, 769, 770, 771, 4096, 5120, 5124, 5125, 5126,
                                         8192, 8199, 16384, 20480])
@pytest.mark.parametrize("add_residual", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
def test_rms_norm(ops, num_tokens, hidden_size, add_residual, dtype):
    g = torch.Generator(device=DEV).manual_seed(0)
    w = (torch.randn(hidden_size, generator=g, device=DEV) * 0.1 + 1.0).to(dtype)
    scale = 1 / (2 * hidden_size)
    x = (torch.randn(num_tokens, hidden_size, generator=g, device=DEV) * scale).to(dtype)
    if add_residual:
        r = (torch.randn(num_tokens, hidden_size, generator=g, device=DEV) * scale).to(dtype)
        ref_x, ref_r = K.fused_add_rms_norm(x, r, w, 1e-6)
        ops.fused_add_rms_norm(x, r, w, 1e-6)
        assert torch.equal(r, ref_r)                    # the residual add is exact in T
        out, ref = x, ref_x
    else:
        ref = K.rms_norm(x, w, 1e-6)
        out = torch.empty_like(x)
        ops.rms_norm(out, x, w, 1e-6)
    assert torch.allclose(out, ref, atol=1e-2, rtol=1e-2)   # the reference test's band
    if dtype != torch.float:
        # only the fp32 scale (sum-of-squares association, rsqrtf) differs from the oracle's
        # float64 one: the first rounding (T)(x*s) can move by one step, the weight multiply can
        # carry that to two; and it must be rare
        assert util.ulp_diff(out, ref) <= 2
        assert (out != ref).float().mean().item() < 2e-2
    else:
        assert torch.allclose(out, ref, atol=1e-6, rtol=1e-5)


def test_rms_norm_large_values(ops):
    """rows whose scaled values are O(1): rounding order (T)(x*s) * w matters here."""
    g = torch.Generator(device=DEV).manual_seed(1)
    for dtype in (torch.half, torch.bfloat16):
        x = (torch.randn(256, 4096, generator=g, device=DEV) * 3).to(dtype)
        r = (torch.randn(256, 4096, generator=g, device=DEV) * 3).to(dtype)
        w = torch.randn(4096, generator=g, device=DEV).to(dtype)
        ref_x, ref_r = K.fused_add_rms_norm(x, r, w, 1e-5)
        ops.fused_add_rms_norm(x, r, w, 1e-5)
        assert torch.equal(r, ref_r)
        assert util.ulp_diff(x, ref_x) <= 2
        assert (x != ref_x).float().mean() < 2e-2


# ---- rotary (tests/kernels/test_pos_encoding.py:23-76) -------------------------------------------
@pytest.mark.parametrize("is_neox", [True, False])
@pytest.mark.parametrize("batch,seq_len", [(1, 11), (5, 11), (1, 8192), (5, 1024)])
@pytest.mark.parametrize("num_heads", [7, 17])
@pytest.mark.parametrize("head_size", [64, 80, 96, 112, 128, 256])
@pytest.mark.parametrize("rotary_dim", [None, 32])
@pytest.mark.parametrize("dtype", DTYPES)
def test_rotary_embedding(ops, is_neox, batch, seq_len, num_heads, head_size, rotary_dim, dtype):
    rot = head_size if rotary_dim is None else rotary_dim
    g = torch.Generator(device=DEV).manual_seed(0)
    cache = util.cos_sin_cache(head_size, rot, 8192, dtype).to(DEV)
    pos = torch.randint(0, 8192, (batch, seq_len), generator=g, device=DEV)
    q = torch.randn(batch, seq_len, num_heads * head_size, generator=g, device=DEV).to(dtype)
    k = torch.randn(batch, seq_len, num_heads * head_size, generator=g, device=DEV).to(dtype)
    rq, rk = K.rotary_embedding(pos, q.reshape(-1, q.shape[-1]), k.reshape(-1, k.shape[-1]),
                                head_size, cache, is_neox)
    ops.rotary_embedding(pos, q, k, head_size, cache, is_neox)
    atol, rtol = TOL[dtype]
    assert torch.allclose(q.reshape(rq.shape), rq, atol=atol, rtol=rtol)
    assert torch.allclose(k.reshape(rk.shape), rk, atol=atol, rtol=rtol)
    if dtype != torch.float:
        assert torch.equal(q.reshape(rq.shape), rq) and torch.equal(k.reshape(rk.shape), rk)


@pytest.mark.parametrize("hq,hkv", [(32, 32), (32, 8), (8, 1)])
@pytest.mark.parametrize("num_tokens", [1, 256, 768])
def test_rotary_embedding_engine_shapes(ops, hq, hkv, num_tokens):
    """q and k are column slices of the fused qkv GEMM output (llama.py:163-165): strided rows,
    GQA (fewer k heads than q heads)."""
    D = 128
    dtype = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(2)
    cache = util.cos_sin_cache(D, D, 8192, dtype).to(DEV)
    pos = torch.randint(0, 8192, (num_tokens,), generator=g, device=DEV)
    qkv = torch.randn(num_tokens, (hq + 2 * hkv) * D, generator=g, device=DEV).to(dtype)
    q, k, v = qkv.split([hq * D, hkv * D, hkv * D], dim=-1)
    rq, rk = K.rotary_embedding(pos, q.contiguous(), k.contiguous(), D, cache, True)
    v0 = v.clone()
    ops.rotary_embedding(pos, q, k, D, cache, True)
    assert torch.equal(q, rq) and torch.equal(k, rk) and torch.equal(v, v0)


@pytest.mark.parametrize("is_neox", [True, False])
@pytest.mark.parametrize("dtype", DTYPES)
def test_batched_rotary_embedding(ops, is_neox, dtype):
    """csrc/pos_encoding_kernels.cu:99-120: per-token offsets into a concatenated cache."""
    D, H, T = 128, 9, 333
    g = torch.Generator(device=DEV).manual_seed(3)
    cache = torch.cat([util.cos_sin_cache(D, D, 4096, dtype, base=b) for b in (10000.0, 50000.0)],
                      dim=0).to(DEV)
    pos = torch.randint(0, 4096, (T,), generator=g, device=DEV)
    offs = torch.randint(0, 2, (T,), generator=g, device=DEV) * 4096
    q = torch.randn(T, H * D, generator=g, device=DEV).to(dtype)
    k = torch.randn(T, H * D, generator=g, device=DEV).to(dtype)
    rq, rk = K.rotary_embedding(pos, q, k, D, cache, is_neox, offsets=offs, rot_dim=D)
    ops.batched_rotary_embedding(pos, q, k, D, cache, is_neox, D, offs)
    atol, rtol = TOL[dtype]
    assert torch.allclose(q, rq, atol=atol, rtol=rtol) and torch.allclose(k, rk, atol=atol, rtol=rtol)


# ---- activations (tests/kernels/test_activation.py:20-86) ----------------------------------------
@pytest.mark.parametrize("act", ["silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"])
@pytest.mark.parametrize("num_tokens", [1, 7, 83, 2048])
@pytest.mark.parametrize("d", [12, 512, 3584, 4096, 5120, 11008, 13824, 14336])
@pytest.mark.parametrize("dtype", DTYPES)
def test_act_and_mul(ops, act, num_tokens, d, dtype):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(num_tokens, 2 * d, generator=g, device=DEV).to(dtype)
    out = torch.empty(num_tokens, d, dtype=dtype, device=DEV)
    getattr(ops, act)(out, x)
    # the reference's own test oracle (layers/activation.py:22-29,51-57), exact: atol = rtol = 0
    if act == "silu_and_mul":
        tref = F.silu(x[..., :d]) * x[..., d:]
    else:
        tref = F.gelu(x[..., :d], approximate="none" if act == "gelu_and_mul" else "tanh") * x[..., d:]
    ref = K.activation(x, act)
    if act == "silu_and_mul":
        assert torch.equal(out, tref)
    else:
        assert torch.allclose(out, tref, atol=TOL[dtype][0], rtol=TOL[dtype][1])
    if dtype != torch.float:
        assert util.ulp_diff(out, ref) <= (0 if act == "silu_and_mul" else 1)
    else:
        assert torch.allclose(out, ref, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("act", ["gelu_new", "gelu_fast"])
@pytest.mark.parametrize("num_tokens", [7, 83, 2048])
@pytest.mark.parametrize("d", [512, 4096, 5120, 13824])
@pytest.mark.parametrize("dtype", DTYPES)
def test_gelu_elementwise(ops, act, num_tokens, d, dtype):
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(num_tokens, d, generator=g, device=DEV).to(dtype)
    out = torch.empty_like(x)
    getattr(ops, act)(out, x)
    xf = x.float()
    if act == "gelu_new":
        tref = 0.5 * xf * (1.0 + torch.tanh(0.79788456 * (xf + 0.044715 * xf ** 3)))
    else:
        tref = 0.5 * xf * (1.0 + torch.tanh(xf * 0.7978845608 * (1.0 + 0.044715 * xf * xf)))
    atol, rtol = TOL[dtype]
    assert torch.allclose(out.float(), tref, atol=max(atol, 4e-3 if dtype != torch.float else 0),
                          rtol=rtol)
    ref = K.activation(x, act)
    if dtype != torch.float:
        assert util.ulp_diff(out, ref) <= 1
        assert (out != ref).float().mean() < 1e-2


def test_elementwise_match_reference_cuda_build(ops, ref_cuda):
    """Bit-exact against the reference's own CUDA kernels (recompiled for sm_100a) for every
    op whose arithmetic has no reduction: silu/gelu family and rotary, f16 + bf16 + f32; rms_norm
    within one rounding step."""
    g = torch.Generator(device=DEV).manual_seed(5)

    def same(a, b):
        # f32: the two builds may contract a*b+c into FMA differently; 16-bit types round the
        # same fp32 intermediates and must agree bit for bit
        if a.dtype == torch.float:
            return torch.allclose(a, b, atol=1e-6, rtol=1e-6)
        return torch.equal(a, b)

    for dtype in DTYPES:
        x = (torch.randn(83, 2 * 5120, generator=g, device=DEV) * 2).to(dtype)
        for name in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
            a = torch.empty(83, 5120, dtype=dtype, device=DEV)
            b = torch.empty_like(a)
            getattr(ops, name)(a, x)
            getattr(ref_cuda.ops, name)(b, x)
            assert same(a, b), (name, dtype)
        for name in ("gelu_new", "gelu_fast"):
            a = torch.empty_like(x)
            b = torch.empty_like(x)
            getattr(ops, name)(a, x)
            getattr(ref_cuda.ops, name)(b, x)
            assert same(a, b), (name, dtype)
        for is_neox in (True, False):
            cache = util.cos_sin_cache(128, 128, 8192, dtype).to(DEV)
            pos = torch.randint(0, 8192, (300,), generator=g, device=DEV)
            q = torch.randn(300, 32 * 128, generator=g, device=DEV).to(dtype)
            k = torch.randn(300, 8 * 128, generator=g, device=DEV).to(dtype)
            q2, k2 = q.clone(), k.clone()
            ops.rotary_embedding(pos, q, k, 128, cache, is_neox)
            ref_cuda.ops.rotary_embedding(pos, q2, k2, 128, cache, is_neox)
            assert same(q, q2) and same(k, k2), (is_neox, dtype)
        w = torch.randn(4096, generator=g, device=DEV).to(dtype)
        h = torch.randn(300, 4096, generator=g, device=DEV).to(dtype)
        r = torch.randn(300, 4096, generator=g, device=DEV).to(dtype)
        h2, r2 = h.clone(), r.clone()
        ops.fused_add_rms_norm(h, r, w, 1e-5)
        ref_cuda.ops.fused_add_rms_norm(h2, r2, w, 1e-5)
        assert torch.equal(r, r2)
        if dtype == torch.float:
            assert torch.allclose(h, h2, atol=1e-5, rtol=1e-5)
        else:
            assert util.ulp_diff(h, h2) <= 2 and (h != h2).float().mean().item() < 2e-2


def test_no_cpu_fallback(ops):
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.silu_and_mul(torch.empty(2, 4), torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match="Unsupported data type"):
        ops.rms_norm(torch.empty(2, 8, dtype=torch.float64, device=DEV),
                     torch.zeros(2, 8, dtype=torch.float64, device=DEV),
                     torch.ones(8, dtype=torch.float64, device=DEV), 1e-5)
