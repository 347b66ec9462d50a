"""GPU parity of the lm_head logits (``vltr_lm_head_logits``, csrc/opt_predictor.cu) against a
float32 torch matmul of the same f16 / bf16 operands - the reference is ``torch.matmul(
hidden_states, embedding.t())`` (logits_processor.py:63), a cuBLAS GEMM with fp32 accumulation
and one rounding.  Stated tolerance: one rounding of the result in the operand type plus
fp32-accumulation noise; the arg-max (what greedy sampling reads) must agree wherever the
float32 top-2 gap exceeds that tolerance."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("rows,vocab,hidden", [
    (256, 32000, 4096),      # Llama-2-7B, a full decode step
    (37, 32000, 4096),
    (256, 4000, 8192),       # Llama-2-70B vocabulary shard at TP8: partial last column tile
    (64, 16032, 4096),       # Llama-3-8B shard at TP8
    (1, 128256, 4096),       # Llama-3-8B, one sequence
    (300, 1000, 256),
])
def test_lm_head_logits(dtype, rows, vocab, hidden):
    from vllm_ltr_b200.logits import lm_head_logits
    g = torch.Generator(device=DEV).manual_seed(rows + vocab)
    h = torch.randn(rows, hidden, generator=g, device=DEV).to(dtype)
    w = (torch.randn(vocab, hidden, generator=g, device=DEV) * 0.02).to(dtype)
    b = torch.randn(vocab, generator=g, device=DEV).to(dtype)
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    for bias in (None, b):
        out = lm_head_logits(h, w, bias)
        assert out.shape == (rows, vocab) and out.dtype == dtype
        ref = h.float() @ w.float().T
        if bias is not None:
            ref = ref + bias.float()
        tol = eps * ref.abs() + 1e-3
        err = (out.float() - ref).abs()
        assert (err <= tol).all(), (err.max().item(), (err / tol).max().item())
        # greedy choice: identical wherever the float32 winner leads by more than the tolerance
        top2 = ref.topk(2, dim=1)
        clear = (top2.values[:, 0] - top2.values[:, 1]) > 4 * eps * top2.values[:, 0].abs() + 4e-3
        assert torch.equal(out.float().argmax(1)[clear], top2.indices[:, 0][clear])


def test_lm_head_logits_against_cublas():
    """Side by side with the reference's own expression on the same device."""
    from vllm_ltr_b200.logits import lm_head_logits
    g = torch.Generator(device=DEV).manual_seed(3)
    h = torch.randn(256, 4096, generator=g, device=DEV).half()
    w = (torch.randn(32000, 4096, generator=g, device=DEV) * 0.02).half()
    ours = lm_head_logits(h, w)
    ref = torch.matmul(h, w.t())
    # both are one fp16 rounding of an fp32-accumulated sum: at most 1 ulp apart, almost all equal
    ulp = (ref.float().abs() * 2.0 ** -10).clamp_min(2.0 ** -14)
    assert ((ours.float() - ref.float()).abs() <= 1.01 * ulp).all()
    assert (ours == ref).float().mean().item() > 0.97


def test_lm_head_logits_errors():
    from vllm_ltr_b200.logits import lm_head_logits
    h = torch.zeros(4, 64, dtype=torch.float16, device=DEV)
    with pytest.raises(RuntimeError, match="no CPU path"):
        lm_head_logits(h.cpu(), torch.zeros(64, 64, dtype=torch.float16))
    with pytest.raises(RuntimeError, match="Unsupported data type"):
        lm_head_logits(h.float(), torch.zeros(64, 64, device=DEV))
    with pytest.raises(RuntimeError, match="multiples of 8"):
        lm_head_logits(h, torch.zeros(100, 64, dtype=torch.float16, device=DEV))
    assert lm_head_logits(h[:0], torch.zeros(64, 64, dtype=torch.float16, device=DEV)).shape == (0, 64)
