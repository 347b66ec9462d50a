"""oracle/kernels.py — CPU restatement of the reference's hot-path kernels.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this
module, and only as the checker.  The product (vllm_ltr_b200) never does.

Each function restates the algorithm of one reference CUDA op, citing the file:line it follows,
with torch-on-CPU / numpy array arithmetic (no custom kernels).  Where the reference computes
"in T" (c10::Half / c10::BFloat16 operators: float op, then round-to-nearest-even back to T) the
restatement rounds at the same places, so for f16/bf16 the element-wise ops are bit-exact
restatements; reductions (sum of squares, softmax, dot products) are evaluated in float64 and
the comparison tolerance covers the association order.

Pinning (see tests/golden/ and DESIGN.md §oracle):
  * against the reference's own CPU kernels (csrc/cpu/*.cpp compiled to oracle/_ref/ref_cpu_C,
    f32/bf16, block 16) on seeded inputs -> tests/golden/cpu_*.npz;
  * against the reference's own CUDA kernels recompiled for sm_100a (oracle/_ref/ref_cuda_C) run
    on a B200 -> tests/golden/cuda_*.npz;
  * against the torch formulas the reference's tests use (tests/kernels/*.py).
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------------------------
# rounding helpers
# --------------------------------------------------------------------------------------------
def rt(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """float32 value of x after a round trip through `dtype` (what "(scalar_t)(expr)" holds)."""
    return x.to(torch.float32).to(dtype).to(torch.float32)


# --------------------------------------------------------------------------------------------
# fp8 e5m2, unscaled  (csrc/quantization/fp8_e5m2_kvcache/quant_utils.cuh:24-30,164-192)
# --------------------------------------------------------------------------------------------
def fp8_e5m2_from(x: torch.Tensor) -> torch.Tensor:
    """__nv_cvt_{float,halfraw,bfloat16raw}_to_fp8(x, __NV_SATFINITE, __NV_E5M2) -> uint8.

    Round-to-nearest-even on the e5m2 grid, computed from the exact value (float64), saturating
    finite overflow to +-57344 (0x7B); NaN -> 0x7F|sign; inf saturates too (SATFINITE)."""
    dev = x.device
    v = x.detach().to("cpu", torch.float64).numpy()
    out = np.zeros(v.shape, dtype=np.uint8)
    sign = (np.signbit(v)).astype(np.uint8) << 7
    a = np.abs(v)
    nan = np.isnan(v)
    # e5m2: 5 exponent bits (bias 15), 2 mantissa bits; min subnormal 2^-16, max 57344
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.floor(np.log2(np.where(a > 0, a, 1.0)))
    e = np.clip(e, -14, 15)              # subnormals share exponent -14
    ulp = np.exp2(e - 2)
    q = a / ulp                          # exact: power-of-two scaling
    r = np.rint(q)                       # numpy rint = round half to even
    val = r * ulp
    val = np.where(np.isinf(a) | (val > 57344.0), 57344.0, val)
    # encode
    with np.errstate(divide="ignore", invalid="ignore"):
        e2 = np.floor(np.log2(np.where(val > 0, val, 1.0)))
    normal = val >= 2.0 ** -14
    e2 = np.where(normal, e2, -14)
    man = np.where(normal, val / np.exp2(e2) - 1.0, val / np.exp2(-14.0)) * 4.0
    code = np.where(normal, (e2 + 15).astype(np.int64) << 2, 0) + man.astype(np.int64)
    out = (code.astype(np.uint8) | sign).astype(np.uint8)
    out = np.where(nan, (0x7F | sign).astype(np.uint8), out)
    return torch.from_numpy(out.astype(np.uint8)).to(dev)


def fp8_e5m2_from_fast(x: torch.Tensor) -> torch.Tensor:
    """Same conversion through torch's float8_e5m2 cast (RNE, but overflow -> inf/NaN), made
    saturating by clamping finite values to +-57344 first.  Device-agnostic; pinned against
    fp8_e5m2_from over all 65536 half bit patterns in tests/test_oracle.py."""
    f = x.to(torch.float32)
    f = torch.where(torch.isnan(f), f, f.clamp(-57344.0, 57344.0))
    # round directly from the source type when it is narrower than float (no double rounding:
    # half/bf16 -> float is exact)
    return f.to(torch.float8_e5m2).view(torch.uint8)


def fp8_e5m2_to(b: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """fp8 -> half is (b << 8) reinterpreted (quant_utils.cuh:24-30); other types go through half
    (:71-82 for bf16, :116-124 for float)."""
    h = (b.to(torch.int32) << 8).to(torch.int16).view(torch.float16)
    return h.to(torch.float32).to(dtype)


# --------------------------------------------------------------------------------------------
# activations  (csrc/activation_kernels.cu:13-52, 132-144)
# --------------------------------------------------------------------------------------------
def _silu(f, T):
    return rt(f / (1.0 + torch.exp(-f)), T)


def activation(x: torch.Tensor, act: str) -> torch.Tensor:
    """silu_and_mul / gelu_and_mul / gelu_tanh_and_mul: out = ACT(x[..., :d]) * x[..., d:], both the
    activation result and the product rounded to T.  gelu_new / gelu_fast: element-wise, every
    intermediate product/sum in T."""
    T = x.dtype
    if act in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
        d = x.shape[-1] // 2
        f = x[..., :d].to(torch.float32)
        y = x[..., d:].to(torch.float32)
        if act == "silu_and_mul":
            a = _silu(f, T)
        elif act == "gelu_and_mul":
            a = rt(f * 0.5 * (1.0 + torch.erf(f * np.float32(0.7071067811865476))), T)
        else:
            beta = np.float32(np.sqrt(2.0) * (2.0 / np.sqrt(np.pi)) * 0.5)
            inner = beta * (f + np.float32(0.044715) * (f * f * f))
            a = rt(0.5 * f * (1.0 + torch.tanh(inner)), T)
        return (a * y).to(T)
    f = x.to(torch.float32)
    c1, c2 = np.float32(0.79788456), np.float32(0.044715)
    if act == "gelu_new":
        x3 = rt(rt(f * f, T) * f, T)
        t = rt(torch.tanh(rt(c1 * rt(f + rt(c2 * x3, T), T), T)), T)
    elif act == "gelu_fast":
        a = rt(f * c1, T)
        b = rt(c2 * f, T)
        t = rt(torch.tanh(rt(a * rt(1.0 + rt(b * f, T), T), T)), T)
    else:
        raise ValueError(act)
    return (rt(rt(torch.tensor(0.5), T) * f, T) * rt(1.0 + t, T)).to(T)


# --------------------------------------------------------------------------------------------
# layernorm  (csrc/layernorm_kernels.cu:22-46, 187-273)
# --------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """out = ((T)(x * rsqrt(mean(x^2) + eps))) * w   — rounds to T before the weight multiply."""
    T = x.dtype
    f = x.to(torch.float32)
    var = (f.to(torch.float64) ** 2).sum(-1, keepdim=True) / x.shape[-1]
    s = (1.0 / torch.sqrt(var + eps)).to(torch.float32)
    return (rt(f * s, T) * w.to(torch.float32)).to(T)


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, w: torch.Tensor,
                       eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """residual' = (T)(x + residual); x' = ((T)(residual' * s)) * w.  Returns (x', residual')."""
    T = x.dtype
    z = (x.to(torch.float32) + residual.to(torch.float32)).to(T)
    return rms_norm(z, w, eps), z


# --------------------------------------------------------------------------------------------
# rotary embedding  (csrc/pos_encoding_kernels.cu:10-97)
# --------------------------------------------------------------------------------------------
def rotary_embedding(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                     head_size: int, cos_sin_cache: torch.Tensor, is_neox: bool,
                     offsets: Optional[torch.Tensor] = None, rot_dim: Optional[int] = None):
    """Returns rotated (query, key) of the input shapes [num_tokens, heads*head_size].  Arithmetic
    in T: x' = (T)((T)(x*cos) - (T)(y*sin)), y' = (T)((T)(y*cos) + (T)(x*sin))."""
    T = query.dtype
    rot = cos_sin_cache.shape[1] if rot_dim is None else rot_dim
    e = rot // 2
    pos = positions.reshape(-1).to(torch.int64)
    if offsets is not None:
        pos = pos + offsets.reshape(-1).to(torch.int64)
    cs = cos_sin_cache[pos].to(torch.float32)       # [tokens, rot]
    cos, sin = cs[:, None, :e], cs[:, None, e:rot]

    def one(arr: torch.Tensor) -> torch.Tensor:
        n = arr.shape[0]
        a = arr.reshape(n, -1, head_size).clone()
        f = a.to(torch.float32)
        if is_neox:
            x, y = f[..., :e], f[..., e:rot]
        else:
            x, y = f[..., 0:rot:2], f[..., 1:rot:2]
        if T == torch.float32:
            xn, yn = x * cos - y * sin, y * cos + x * sin
        else:
            xn = rt(rt(x * cos, T) - rt(y * sin, T), T)
            yn = rt(rt(y * cos, T) + rt(x * sin, T), T)
        if is_neox:
            a[..., :e], a[..., e:rot] = xn.to(T), yn.to(T)
        else:
            a[..., 0:rot:2], a[..., 1:rot:2] = xn.to(T), yn.to(T)
        return a.reshape(arr.shape)

    return one(query), one(key)


# --------------------------------------------------------------------------------------------
# cache ops  (csrc/cache_kernels.cu:68-93, 157-216)
# --------------------------------------------------------------------------------------------
def reshape_and_cache(key: torch.Tensor, value: torch.Tensor, key_cache: torch.Tensor,
                      value_cache: torch.Tensor, slot_mapping: torch.Tensor,
                      kv_cache_dtype: str = "auto") -> None:
    """In place on CPU copies of the caches.  key/value [T, H, D]; key_cache [nb, H, D/x, bs, x];
    value_cache [nb, H, D, bs]; slot < 0 is skipped (:172-175)."""
    nb, H, Dx, bs, x = key_cache.shape
    slots = slot_mapping.reshape(-1).to(torch.int64)
    keep = slots >= 0
    t_idx = torch.nonzero(keep).reshape(-1)
    blk = torch.div(slots[keep], bs, rounding_mode="floor")
    off = slots[keep] % bs
    k = key[t_idx]
    v = value[t_idx]
    if kv_cache_dtype == "fp8":
        k, v = fp8_e5m2_from(k), fp8_e5m2_from(v)
    D = Dx * x
    # tokens are applied in order so a later duplicate slot wins, like sequential CTAs would not
    # guarantee — the engine never issues duplicate slots; tests do not either.
    key_cache[blk, :, :, off, :] = k.reshape(-1, H, Dx, x)
    value_cache[blk, :, :, off] = v.reshape(-1, H, D)


def copy_blocks(key_caches: List[torch.Tensor], value_caches: List[torch.Tensor],
                block_mapping: Dict[int, List[int]]) -> None:
    """cache_kernels.cu:68-93: for every layer, K and V block src -> each dst."""
    for src, dsts in block_mapping.items():
        for dst in dsts:
            for kc in key_caches:
                kc[dst].copy_(kc[src])
            for vc in value_caches:
                vc[dst].copy_(vc[src])


def swap_blocks(src: torch.Tensor, dst: torch.Tensor, block_mapping: Dict[int, int]) -> None:
    """cache_kernels.cu:23-62: dst[d] = src[s] block-wise."""
    for s, d in block_mapping.items():
        dst[d].copy_(src[s])


# --------------------------------------------------------------------------------------------
# paged attention  (csrc/attention/attention_kernels.cu:97-442; the reference test's oracle
# tests/kernels/test_attention.py:57-113 is the same computation)
# --------------------------------------------------------------------------------------------
def paged_attention(query: torch.Tensor, key_cache: torch.Tensor, value_cache: torch.Tensor,
                    block_tables: torch.Tensor, context_lens: torch.Tensor, scale: float,
                    alibi_slopes: Optional[torch.Tensor] = None,
                    kv_cache_dtype: str = "auto") -> torch.Tensor:
    """out[i,h] = softmax(scale*q[i,h].K_i^T + alibi) . V_i over the paged KV, float64 math.
    softmax denominator carries the reference's +1e-6 (:293).  Returns float32 [B, Hq, D]."""
    B, Hq, D = query.shape
    nb, Hkv, Dx, bs, x = key_cache.shape
    G = Hq // Hkv
    dev = query.device
    out = torch.zeros(B, Hq, D, dtype=torch.float32, device=dev)
    context_lens = context_lens.to("cpu")
    if kv_cache_dtype == "fp8":
        key_cache = fp8_e5m2_to(key_cache, torch.float32)
        value_cache = fp8_e5m2_to(value_cache, torch.float32)
    q64 = query.to(torch.float64)
    for i in range(B):
        L = int(context_lens[i])
        if L == 0:
            continue
        nblk = (L + bs - 1) // bs
        ids = block_tables[i, :nblk].to(torch.int64)
        # K [nblk, Hkv, D/x, bs, x] -> [nblk*bs, Hkv, D]
        k = key_cache[ids].permute(0, 3, 1, 2, 4).reshape(nblk * bs, Hkv, D)[:L].to(torch.float64)
        v = value_cache[ids].permute(0, 3, 1, 2).reshape(nblk * bs, Hkv, D)[:L].to(torch.float64)
        k = k.repeat_interleave(G, dim=1)      # kv head of q head h is h // G (:148-149)
        v = v.repeat_interleave(G, dim=1)
        logits = scale * torch.einsum("hd,lhd->hl", q64[i], k)
        if alibi_slopes is not None:
            pos = torch.arange(L, dtype=torch.float64, device=dev) - L + 1     # (:248)
            logits = logits + alibi_slopes.to(torch.float64)[:, None] * pos[None, :]
        m = logits.max(dim=1, keepdim=True).values
        e = torch.exp(logits - m)
        p = e / (e.sum(dim=1, keepdim=True) + 1e-6)
        out[i] = torch.einsum("hl,lhd->hd", p, v).to(torch.float32)
    return out


def paged_attention_partitions(query: torch.Tensor, key_cache: torch.Tensor,
                               value_cache: torch.Tensor, block_tables: torch.Tensor,
                               context_lens: torch.Tensor, scale: float, partition_size: int = 512,
                               alibi_slopes: Optional[torch.Tensor] = None,
                               kv_cache_dtype: str = "auto"):
    """What paged_attention_v2_kernel leaves in its side outputs (:382-441): for every partition
    j of 512 tokens that intersects the context, max_logits[i,h,j] = max of the (scaled, biased)
    logits, exp_sums[i,h,j] = sum exp(logit - max), tmp_out[i,h,j] = softmax-weighted V with the
    partition-local denominator (exp_sum + 1e-6).  Entries of partitions past the context are
    NaN here (the reference never writes them).  Returns float32 (max_logits, exp_sums, tmp_out)."""
    B, Hq, D = query.shape
    nb, Hkv, Dx, bs, x = key_cache.shape
    G = Hq // Hkv
    dev = query.device
    context_lens = context_lens.to("cpu")
    max_ctx = int(context_lens.max()) if B else 0
    P = max(1, (max_ctx + partition_size - 1) // partition_size)
    ml = torch.full((B, Hq, P), float("nan"), dtype=torch.float32, device=dev)
    es = torch.full((B, Hq, P), float("nan"), dtype=torch.float32, device=dev)
    to = torch.full((B, Hq, P, D), float("nan"), dtype=torch.float32, device=dev)
    if kv_cache_dtype == "fp8":
        key_cache = fp8_e5m2_to(key_cache, torch.float32)
        value_cache = fp8_e5m2_to(value_cache, torch.float32)
    q64 = query.to(torch.float64)
    for i in range(B):
        L = int(context_lens[i])
        if L == 0:
            continue
        nblk = (L + bs - 1) // bs
        ids = block_tables[i, :nblk].to(torch.int64)
        k = key_cache[ids].permute(0, 3, 1, 2, 4).reshape(nblk * bs, Hkv, D)[:L].to(torch.float64)
        v = value_cache[ids].permute(0, 3, 1, 2).reshape(nblk * bs, Hkv, D)[:L].to(torch.float64)
        k = k.repeat_interleave(G, dim=1)
        v = v.repeat_interleave(G, dim=1)
        logits = scale * torch.einsum("hd,lhd->hl", q64[i], k)
        if alibi_slopes is not None:
            pos = torch.arange(L, dtype=torch.float64, device=dev) - L + 1
            logits = logits + alibi_slopes.to(torch.float64)[:, None] * pos[None, :]
        for j in range((L + partition_size - 1) // partition_size):
            lo, hi = j * partition_size, min(L, (j + 1) * partition_size)
            lg = logits[:, lo:hi]
            m = lg.max(dim=1, keepdim=True).values
            e = torch.exp(lg - m)
            s = e.sum(dim=1, keepdim=True)
            ml[i, :, j] = m[:, 0].to(torch.float32)
            es[i, :, j] = s[:, 0].to(torch.float32)
            to[i, :, j] = torch.einsum("hl,lhd->hd", e / (s + 1e-6), v[lo:hi]).to(torch.float32)
    return ml, es, to


def context_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_cache: torch.Tensor,
                      value_cache: torch.Tensor, block_tables: torch.Tensor,
                      start_loc: torch.Tensor, seq_lens: torch.Tensor, ctx_lens: torch.Tensor,
                      alibi_slopes: Optional[torch.Tensor] = None,
                      kv_cache_dtype: str = "auto") -> torch.Tensor:
    """Prefill attention with a cached prefix (vllm/attention/ops/prefix_prefill.py:10-628; the
    reference's own test oracle, tests/kernels/test_prefix_prefill.py:60-170, is the same maths
    with xformers' BlockDiagonalCausalFromBottomRightMask): token i of sequence b attends to the
    first ctx_b tokens of the paged cache and to chunk tokens 0..i taken from k / v.  float64;
    returns float32 [num_tokens, Hq, D]."""
    T, Hq, D = q.shape
    Hkv = k.shape[1]
    G = Hq // Hkv
    nb, _, Dx, bs, x = key_cache.shape
    dev = q.device
    out = torch.zeros(T, Hq, D, dtype=torch.float32, device=dev)
    if kv_cache_dtype == "fp8":
        key_cache = fp8_e5m2_to(key_cache, torch.float32)
        value_cache = fp8_e5m2_to(value_cache, torch.float32)
    scale = D ** -0.5
    for b in range(seq_lens.shape[0]):
        c, n = int(ctx_lens[b]), int(seq_lens[b]) - int(ctx_lens[b])
        if n <= 0:
            continue
        s0 = int(start_loc[b])
        nblk = (c + bs - 1) // bs
        ids = block_tables[b, :nblk].to(torch.int64)
        kp = key_cache[ids].permute(0, 3, 1, 2, 4).reshape(nblk * bs, Hkv, D)[:c].to(torch.float64)
        vp = value_cache[ids].permute(0, 3, 1, 2).reshape(nblk * bs, Hkv, D)[:c].to(torch.float64)
        kk = torch.cat([kp, k[s0:s0 + n].to(torch.float64)]).repeat_interleave(G, dim=1)
        vv = torch.cat([vp, v[s0:s0 + n].to(torch.float64)]).repeat_interleave(G, dim=1)
        logits = scale * torch.einsum("qhd,khd->hqk", q[s0:s0 + n].to(torch.float64), kk)
        qpos = c + torch.arange(n, device=dev)
        kpos = torch.arange(c + n, device=dev)
        if alibi_slopes is not None:
            logits = logits + alibi_slopes.to(torch.float64)[:, None, None] * \
                (kpos[None, None, :] - qpos[None, :, None]).to(torch.float64)
        logits = logits.masked_fill(kpos[None, None, :] > qpos[None, :, None], float("-inf"))
        p = torch.softmax(logits, dim=-1)
        out[s0:s0 + n] = torch.einsum("hqk,khd->qhd", p, vv).to(torch.float32)
    return out


def make_kv_cache(num_blocks: int, block_size: int, num_kv_heads: int, head_size: int,
                  dtype: torch.dtype, seed: int, kv_cache_dtype: str = "auto"):
    """Random caches in the paged layout, uniform(-1/sqrt(D), 1/sqrt(D)) like the reference fixture
    (vllm/utils.py:344-410).  Returns CPU tensors (key_cache, value_cache)."""
    g = torch.Generator().manual_seed(seed)
    scale = head_size ** -0.5
    esz = 1 if kv_cache_dtype == "fp8" else torch.tensor([], dtype=dtype).element_size()
    x = 16 // esz
    kshape = (num_blocks, num_kv_heads, head_size // x, block_size, x)
    vshape = (num_blocks, num_kv_heads, head_size, block_size)
    k = (torch.rand(kshape, generator=g) * 2 - 1) * scale
    v = (torch.rand(vshape, generator=g) * 2 - 1) * scale
    if kv_cache_dtype == "fp8":
        return fp8_e5m2_from(k.to(dtype)), fp8_e5m2_from(v.to(dtype))
    return k.to(dtype), v.to(dtype)
