"""GPU parity of paged_attention_v1/v2 (through the C ABI via the vllm._C mirror) against the
oracle.  Grid and tolerances are the reference's (tests/kernels/test_attention.py:113-274:
atol 1e-3 / rtol 1e-5, fp8 atol 1e-2); edge cases added: ragged / tiny / empty contexts, context
lengths past the reference's shared-memory cap, strided queries, block size 8, GQA group sizes
of the benchmark configs, and the v2 side outputs."""
import pytest
import torch

from tests import util
from oracle import kernels as K

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
DTYPES = [torch.half, torch.bfloat16, torch.float]
NUM_HEADS = [(40, 40), (64, 8)]
HEAD_SIZES = [64, 80, 96, 112, 128, 256]
BLOCK_SIZES = [16, 32]


@pytest.fixture(scope="module")
def ops():
    from vllm_ltr_b200 import _C
    return _C.ops


def _check(out, ref, kv_cache_dtype):
    atol, rtol = (1e-2, 1e-5) if kv_cache_dtype == "fp8" else (1e-3, 1e-5)
    o = out.float()
    assert torch.isfinite(o).all()
    err = (o - ref).abs().max().item()
    assert torch.allclose(o, ref, atol=atol, rtol=rtol), f"max abs err {err}"


@pytest.mark.parametrize("version", ["v1", "v2"])
@pytest.mark.parametrize("num_heads", NUM_HEADS)
@pytest.mark.parametrize("head_size", HEAD_SIZES)
@pytest.mark.parametrize("use_alibi", [False, True])
@pytest.mark.parametrize("block_size", BLOCK_SIZES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kv_cache_dtype", ["auto", "fp8"])
def test_paged_attention_grid(ops, version, num_heads, head_size, use_alibi, block_size, dtype,
                              kv_cache_dtype):
    """The reference grid (7 sequences, random contexts, last one the maximum)."""
    c = util.attention_case(7, num_heads, head_size, block_size, dtype, kv_cache_dtype,
                            use_alibi, 0, 2100, DEV, num_blocks=1024)
    out, _ = util.run_attention(ops, c, version)
    _check(out, util.oracle_attention(c), kv_cache_dtype)


@pytest.mark.parametrize("version", ["v1", "v2"])
@pytest.mark.parametrize("dtype", [torch.half, torch.bfloat16])
@pytest.mark.parametrize("kv_cache_dtype", ["auto", "fp8"])
@pytest.mark.parametrize("max_len", [57 * 1024 - 512, 100_000])
def test_paged_attention_long(ops, version, dtype, kv_cache_dtype, max_len):
    """The reference's maximum (smem/4 - 512 tokens, test_attention.py:16) and beyond it: this
    kernel keeps no per-context state in shared memory, so there is no cap."""
    c = util.attention_case(7, (64, 8), 128, 16, dtype, kv_cache_dtype, False, 1, max_len, DEV)
    out, _ = util.run_attention(ops, c, version)
    _check(out, util.oracle_attention(c), kv_cache_dtype)


@pytest.mark.parametrize("version", ["v1", "v2"])
@pytest.mark.parametrize("num_heads", [(32, 32), (32, 8), (8, 1), (12, 12), (16, 2), (24, 4),
                                       (32, 2), (24, 2), (40, 2)])   # groups of 16, 12, 20 (2 passes)
@pytest.mark.parametrize("block_size", [8, 16, 32])
@pytest.mark.parametrize("dtype", [torch.half, torch.bfloat16])
@pytest.mark.parametrize("kv_cache_dtype", ["auto", "fp8"])
def test_paged_attention_ragged(ops, version, num_heads, block_size, dtype, kv_cache_dtype):
    """Ragged batches around block / partition boundaries, including empty contexts."""
    if kv_cache_dtype == "fp8" and num_heads[0] // num_heads[1] < 8 and num_heads != (32, 8):
        pytest.skip("e5m2 variants: the large groups and one small one")
    lens = [1, 2, block_size - 1, block_size, block_size + 1, 511, 512, 513, 1023, 1025, 0, 31,
            777, 3, 2048, 1]
    c = util.attention_case(len(lens), num_heads, 128, block_size, dtype, kv_cache_dtype, False,
                            3, 0, DEV, context_lens=lens, num_blocks=512)
    out, _ = util.run_attention(ops, c, version)
    ref = util.oracle_attention(c)
    _check(out, ref, kv_cache_dtype)
    assert (out[10] == 0).all()      # empty context: zeros, like the reference's untouched accs


@pytest.mark.parametrize("num_seqs", [1, 3, 64, 256])
@pytest.mark.parametrize("num_heads,dtype,kv", [((32, 32), torch.half, "auto"),
                                                ((32, 8), torch.bfloat16, "auto"),
                                                ((32, 8), torch.bfloat16, "fp8"),
                                                ((8, 1), torch.bfloat16, "auto"),
                                                ((4, 4), torch.half, "auto"),
                                                ((4, 1), torch.bfloat16, "fp8")])
@pytest.mark.parametrize("split", [0, 1, -1])
def test_paged_attention_bench_shapes(ops, num_seqs, num_heads, dtype, kv, split):
    """The micro-shapes of SURVEY.md section 8d: batch 1..256, the (Hq, Hkv, dtype) configs of the
    benchmark models and of their TP8 shards, ShareGPT-like contexts; batch 1 and 3 exercise the
    library-side context split of v1.  `split`: the library's own choice between the one-launch
    form (partitions merged by the last finisher) and the two-launch form of the latency-bound
    regime (kernel-sized partitions + parallel merge kernel), then each of them forced."""
    from vllm_ltr_b200._lib import tune
    c = util.attention_case(num_seqs, num_heads, 128, 16, dtype, kv, False, 7,
                            6000 if num_seqs <= 3 else 1500, DEV, num_blocks=8192)
    tune("attn_split", split)
    try:
        for version in ("v1", "v2"):
            out, _ = util.run_attention(ops, c, version)
            _check(out, util.oracle_attention(c), kv)
    finally:
        tune("attn_split", 0)


@pytest.mark.parametrize("min_tokens", [16, 64, 256])
@pytest.mark.parametrize("num_heads", [(8, 1), (4, 4), (40, 2)])
def test_paged_attention_split_partition_sizes(ops, min_tokens, num_heads):
    """Two-launch form with small / large minimum partitions: ragged contexts around partition
    edges, an empty one, a long one (many partitions for the merge kernel)."""
    from vllm_ltr_b200._lib import tune
    lens = [1, 15, 16, 17, 63, 64, 65, 127, 129, 0, 255, 257, 1000, 4097, 33, 8000]
    c = util.attention_case(len(lens), num_heads, 128, 16, torch.bfloat16, "auto", True, 13, 0,
                            DEV, context_lens=lens, num_blocks=1024)
    tune("attn_split", 1)
    tune("attn_min_tokens", min_tokens)
    try:
        out, _ = util.run_attention(ops, c, "v1")
    finally:
        tune("attn_split", 0)
        tune("attn_min_tokens", 0)
    _check(out, util.oracle_attention(c), "auto")
    assert (out[9] == 0).all()


def test_paged_attention_strided_query(ops):
    """query is a strided view of the fused qkv projection (attention_kernels.cu:175-176)."""
    c = util.attention_case(9, (32, 8), 128, 16, torch.bfloat16, "auto", False, 5, 900, DEV,
                            num_blocks=512)
    B, H, D = c["query"].shape
    qkv = torch.randn(B, (H + 16) * D, dtype=torch.bfloat16, device=DEV)
    qkv[:, :H * D] = c["query"].reshape(B, H * D)
    c["query"] = qkv[:, :H * D].view(B, H, D)
    assert c["query"].stride(0) == (H + 16) * D
    out = torch.empty(B, H, D, dtype=torch.bfloat16, device=DEV)
    ops.paged_attention_v1(out, c["query"], c["key_cache"], c["value_cache"], 8, c["scale"],
                           c["block_tables"], c["context_lens"], 16, c["max_ctx"], None, "auto",
                           1.0)
    _check(out, util.oracle_attention(c), "auto")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("num_heads", [(16, 16), (32, 8)])
def test_paged_attention_v2_side_outputs(ops, dtype, num_heads):
    """v2 leaves max_logits / exp_sums / tmp_out for every partition that intersects the context,
    with the reference's 512-token partitions (attention_kernels.cu:382-441)."""
    c = util.attention_case(6, num_heads, 128, 16, dtype, "auto", True, 11, 0, DEV,
                            context_lens=[5, 512, 513, 1400, 2047, 1024], num_blocks=512)
    out, (ml, es, tmp) = util.run_attention(ops, c, "v2")
    rml, res, rtmp = K.paged_attention_partitions(c["query"], c["key_cache"], c["value_cache"],
                                                  c["block_tables"], c["context_lens"],
                                                  c["scale"], 512, c["alibi"])
    live = ~torch.isnan(rml)
    assert torch.allclose(ml[live], rml[live], atol=1e-3, rtol=1e-3)
    assert torch.allclose(es[live], res[live], atol=1e-2, rtol=1e-2)   # p is rounded to T
    assert torch.allclose(tmp.float()[live], rtmp[live], atol=2e-3, rtol=1e-2)
    _check(out, util.oracle_attention(c), "auto")


def test_paged_attention_matches_reference_cuda_build(ops, ref_cuda):
    """Side by side with the reference's own kernel recompiled for sm_100a (oracle/_ref): both
    must sit inside the tolerance band around the oracle, and close to each other."""
    for dtype in (torch.half, torch.bfloat16, torch.float):
        for kv in ("auto", "fp8"):
            c = util.attention_case(7, (64, 8), 128, 16, dtype, kv, True, 2, 3000, DEV)
            ours, _ = util.run_attention(ops, c, "v1")
            theirs, _ = util.run_attention(ref_cuda.ops, c, "v1")
            ref = util.oracle_attention(c)
            _check(ours, ref, kv)
            _check(theirs, ref, kv)
            e_ours = (ours.float() - ref).abs().max().item()
            e_theirs = (theirs.float() - ref).abs().max().item()
            assert e_ours <= max(2 * e_theirs, 2e-4), (e_ours, e_theirs)


def test_paged_attention_errors(ops):
    c = util.attention_case(2, (8, 8), 128, 16, torch.half, "auto", False, 0, 64, DEV,
                            num_blocks=16)
    out = torch.empty_like(c["query"])
    with pytest.raises(RuntimeError, match="Unsupported head size"):
        q = torch.zeros(2, 8, 72, dtype=torch.half, device=DEV)
        kc = torch.zeros(16, 8, 9, 16, 8, dtype=torch.half, device=DEV)
        vc = torch.zeros(16, 8, 72, 16, dtype=torch.half, device=DEV)
        ops.paged_attention_v1(torch.empty_like(q), q, kc, vc, 8, 1.0, c["block_tables"],
                               c["context_lens"], 16, 64, None, "auto", 1.0)
    with pytest.raises(RuntimeError, match="Unsupported block size"):
        kc = torch.zeros(16, 8, 16, 4, 8, dtype=torch.half, device=DEV)
        vc = torch.zeros(16, 8, 128, 4, dtype=torch.half, device=DEV)
        ops.paged_attention_v1(out, c["query"], kc, vc, 8, 1.0, c["block_tables"],
                               c["context_lens"], 4, 64, None, "auto", 1.0)
    with pytest.raises(RuntimeError, match="Unsupported data type of kv cache"):
        ops.paged_attention_v1(out, c["query"], c["key_cache"], c["value_cache"], 8, 1.0,
                               c["block_tables"], c["context_lens"], 16, 64, None, "int4", 1.0)
    with pytest.raises(RuntimeError):
        ops.paged_attention_v1(out, c["query"].cpu(), c["key_cache"], c["value_cache"], 8, 1.0,
                               c["block_tables"], c["context_lens"], 16, 64, None, "auto", 1.0)
