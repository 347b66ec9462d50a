"""GPU parity of the LTR predictor (csrc/opt_predictor.cu) through the C ABI.

* the tcgen05 dense layer against a float32 torch matmul of the same fp16/bf16 inputs (the one
  floating-point GEMM in scope keeps a torch fp32 reference): fp32-accumulate tolerance;
* the whole forward against oracle/opt_oracle.py (float64 restatement of the reference's
  OPTForSequenceClassification, pinned against HF transformers on CPU) on seeded random weights of
  the exact OPT-125M / OPT-350M layouts.  Stated tolerance (the reference
  has no predictor test): |logit - oracle| <= 2e-2 for fp16 activations, and the ORDER of the
  scores - what the scheduler consumes - agrees with the oracle's order (Kendall tau >= 0.99).
"""
import pytest
import torch

from oracle import opt_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def P():
    from vllm_ltr_b200 import predictor
    return predictor


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("m", [1, 7, 128, 129, 300, 2048])
@pytest.mark.parametrize("n,k", [(2304, 768), (768, 768), (3072, 768), (768, 3072), (1024, 512),
                                 (512, 1024), (64, 64), (4096, 1024),
                                 # partial last column tile (vocabulary shards of the lm_head)
                                 (4000, 512), (200, 64), (72, 128), (8, 64)])
def test_dense_layer(P, dtype, m, n, k):
    g = torch.Generator(device=DEV).manual_seed(m * 7 + n + k)
    a = torch.randn(m, k, generator=g, device=DEV).to(dtype)
    w = (torch.randn(n, k, generator=g, device=DEV) * 0.05).to(dtype)
    b = torch.randn(n, generator=g, device=DEV).to(dtype)
    r = torch.randn(m, n, generator=g, device=DEV).to(dtype)
    ref = a.float() @ w.float().T
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7

    def close(out, want):
        err = (out.float() - want).abs()
        # one rounding of the result, plus (residual form) one of the linear output before it
        tol = 2 * eps * (want.abs() + ref.abs() + 1) + 1e-3
        assert (err <= tol).all(), (err.max().item(), (err / tol).max().item())

    close(P.gemm(a, w), ref)
    close(P.gemm(a, w, bias=b), ref + b.float())
    close(P.gemm(a, w, bias=b, relu=True), torch.relu(ref + b.float()))
    lin = (ref + b.float()).to(dtype).float()            # linear output is T, then + residual
    close(P.gemm(a, w, bias=b, resid=r), lin + r.float())


def test_dense_layer_errors(P):
    a = torch.zeros(4, 64, dtype=torch.float16, device=DEV)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        P.gemm(a, torch.zeros(100, 64, dtype=torch.float16, device=DEV))
    with pytest.raises(RuntimeError, match="Unsupported data type"):
        P.gemm(a.float(), torch.zeros(64, 64, device=DEV))
    with pytest.raises(RuntimeError, match="no CPU path"):
        P.gemm(a.cpu(), torch.zeros(64, 64, dtype=torch.float16))


def _prompts(n, max_len, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, max_len + 1, (n,), generator=g).tolist()
    lens[0], lens[-1] = 1, max_len
    return [torch.randint(3, vocab, (L,), generator=g).tolist() for L in lens]


def kendall_tau(a, b):
    import itertools
    c = d = 0
    for i, j in itertools.combinations(range(len(a)), 2):
        s = (a[i] - a[j]) * (b[i] - b[j])
        c += s > 0
        d += s < 0
    return (c - d) / max(1, c + d)


@pytest.mark.parametrize("kind,layers", [("125m", 12), ("350m", 24)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_forward_matches_oracle(P, kind, layers, dtype):
    cfg, sd = O.random_opt(kind, seed=11, vocab=2000, max_pos=512, layers=layers)
    pred = P.OPTPredictor(cfg, sd, device=DEV, dtype=dtype)
    prompts = _prompts(24, 300, 2000, 5)
    got = pred.logits(prompts).cpu()
    want = O.opt_logits(cfg, sd, prompts)
    tol = 2e-2 if dtype == torch.float16 else 1e-1
    assert (got - want).abs().max().item() <= tol, (got - want).abs().max().item()
    # fp16 is what the predictor checkpoints are saved in (train/trainer.py:224); bf16 logits
    # are quantised to 2^-8 relative, so near-equal scores legitimately tie or swap
    assert kendall_tau(got[:, 0].tolist(), want[:, 0].tolist()) >= (0.99 if dtype == torch.float16 else 0.85)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_tile_edges(P, dtype):
    """Prompt lengths around the 64-row query / 64-key tile edges of the tensor-core attention
    kernel (and 1-token prompts), in one varlen batch."""
    cfg, sd = O.random_opt("125m", seed=3, vocab=1500, max_pos=1024, layers=2)
    g = torch.Generator().manual_seed(17)
    lens = [1, 2, 15, 16, 17, 63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 500, 1, 1024]
    prompts = [torch.randint(3, 1500, (L,), generator=g).tolist() for L in lens]
    got = P.OPTPredictor(cfg, sd, device=DEV, dtype=dtype).logits(prompts).cpu()
    want = O.opt_logits(cfg, sd, prompts)
    tol = 2e-2 if dtype == torch.float16 else 1e-1
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= tol, (got - want).abs().max().item()


def test_forward_batching_is_invisible(P):
    """Scores do not depend on how prompts are packed into forwards (the reference's aux
    scheduler packs <= 2048 tokens / <= 256 prompts per step)."""
    cfg, sd = O.random_opt("125m", seed=2, vocab=1000, max_pos=2048, layers=2)
    prompts = _prompts(40, 700, 1000, 9)
    a = P.OPTPredictor(cfg, sd, device=DEV, max_batch_tokens=2048).logits(prompts)
    b = P.OPTPredictor(cfg, sd, device=DEV, max_batch_tokens=1).logits(prompts)   # one by one
    c = P.OPTPredictor(cfg, sd, device=DEV).logits(prompts)       # default: one 16k-token batch
    assert torch.equal(a, b) and torch.equal(a, c)
    long = [[3 + (i * 7) % 990 for i in range(2048)]]                              # max_length
    c = P.OPTPredictor(cfg, sd, device=DEV).logits(long).cpu()
    assert (c - O.opt_logits(cfg, sd, long)).abs().max().item() <= 2e-2


def test_classification_head_and_aux_model(P):
    """mtype == "class": the score is the argmax label as float (opt.py:394-395); and the
    scheduler-facing object sets aux_model_score on every group."""
    cfg, sd = O.random_opt("125m", seed=4, num_labels=8, vocab=1000, max_pos=256, layers=2)
    pred = P.OPTPredictor(cfg, sd, device=DEV)
    prompts = _prompts(16, 100, 1000, 1)
    lg = O.opt_logits(cfg, sd, prompts)
    got, want = pred.scores(prompts), lg.argmax(-1).float().tolist()
    top2 = lg.topk(2, dim=-1).values
    clear = ((top2[:, 0] - top2[:, 1]) > 5e-2).tolist()       # skip near-ties of the oracle
    assert all(g == w for g, w, c in zip(got, want, clear) if c) and sum(clear) >= 8

    class Seq:
        def __init__(self, text):
            self.prompt = text

    class Group:
        def __init__(self, text):
            self.seqs_dict = {0: Seq(text)}
            self.aux_model_score = None

        def need_aux_model_score(self):
            return self.aux_model_score is None

        def set_aux_model_score(self, s):
            self.aux_model_score = s

    cfg1, sd1 = O.random_opt("125m", seed=4, vocab=1000, max_pos=64, layers=1)
    tok = lambda text: [2] + [3 + (ord(ch) % 900) for ch in text]
    aux = P.B200AuxModel(P.OPTPredictor(cfg1, sd1, device=DEV), tok, max_length=16)
    groups = [Group("hello world " * k) for k in range(1, 6)]
    ret = aux.obtain_aux_scores(groups)
    want = O.opt_scores(cfg1, sd1, [tok(g.seqs_dict[0].prompt)[:16] for g in groups])
    assert [g.aux_model_score for g in groups] == ret
    assert max(abs(a - b) for a, b in zip(ret, want)) <= 2e-2
    assert all(float(torch.tensor(x, dtype=torch.float16)) == x for x in ret)   # fp16 floats


def test_aux_model_burst_in_chunks(P):
    """A burst of unscored requests goes through obtain_aux_scores in chunks (next chunk tokenised
    on the helper thread while the forwards of the current one are enqueued, staging ring of
    three): every group gets the score of ITS prompt - the same, to fp16 round-off of the model,
    as in one chunk and as the oracle's (stated tolerance 2e-2, see the module docstring)."""
    class Seq:
        def __init__(self, text):
            self.prompt = text

    class Group:
        def __init__(self, text):
            self.seqs_dict = {0: Seq(text)}
            self.aux_model_score = None

        def need_aux_model_score(self):
            return self.aux_model_score is None

        def set_aux_model_score(self, s):
            self.aux_model_score = s

    cfg, sd = O.random_opt("125m", seed=6, vocab=1000, max_pos=64, layers=1)
    tok = lambda text: [2] + [3 + (ord(ch) * 7 + i) % 900 for i, ch in enumerate(text)]
    texts = [("scheduling by rank " * 3)[: 2 + (11 * i) % 47] + chr(97 + i % 26) for i in range(41)]
    pred = P.OPTPredictor(cfg, sd, device=DEV, max_batch_tokens=256)    # several forwards per chunk
    batch = lambda ts: [tok(t) for t in ts]
    whole = P.B200AuxModel(pred, tok, max_length=40, batch_tokenizer=batch, chunk_prompts=1000)
    parts = P.B200AuxModel(pred, tok, max_length=40, batch_tokenizer=batch, chunk_prompts=6)
    ga, gb = [Group(t) for t in texts], [Group(t) for t in texts]
    a, b = whole.obtain_aux_scores(ga), parts.obtain_aux_scores(gb)
    assert [g.aux_model_score for g in gb] == b and len(b) == len(texts)
    want = O.opt_scores(cfg, sd, [tok(t)[:40] for t in texts])
    assert max(abs(x - y) for x, y in zip(a, b)) <= 2e-2
    assert max(abs(x - y) for x, y in zip(b, want)) <= 2e-2
