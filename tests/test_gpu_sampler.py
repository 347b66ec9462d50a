"""GPU: vltr_sample (csrc/sampler.cu) against the reference Sampler's recorded outputs
(tests/golden/sampler_ref.npz), against the oracle on seeded inputs at the real vocabulary sizes,
and - for the in-kernel Philox noise, which no reference run can reproduce - through the
distribution it must follow."""
import pytest
import torch

from oracle import sampler_oracle as O
from tests.test_oracle_sampler import load_cases, split_temps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ours(logits, temps=None, rnd=None, noise=None, **kw):
    from vllm_ltr_b200.sampler import sample_rows
    return sample_rows(logits.to(DEV), temps, rnd, noise, **kw)


def check_against(tok, lp, rk, want_tok, want_lp, want_rk, logits, temps, noise, rnd):
    """tokens / ranks exact, logprob to fp32 round-off.  A random row may legitimately differ when
    the two best keys p/q are closer than fp32 resolves (the kernel ranks x - log q, the reference
    p / q): such a row must then hold the reference's runner-up."""
    tok, lp, rk = tok.cpu(), lp.cpu(), rk.cpu()
    bad = (tok != want_tok).nonzero().flatten().tolist()
    for r in bad:
        assert rnd is not None and bool(rnd[r]), f"greedy row {r} differs"
        x = logits[r].float() / (temps[r] if temps is not None else 1.0)
        key = x - torch.log(noise[r])
        top2 = torch.topk(key, 2)
        assert int(tok[r]) in top2.indices.tolist() and float(top2.values[0] - top2.values[1]) < 1e-5
    ok = tok == want_tok
    assert torch.equal(rk[ok], want_rk[ok])
    assert torch.allclose(lp[ok], want_lp[ok], rtol=1e-5, atol=2e-5)
    assert len(bad) <= max(1, len(tok) // 100)


@pytest.mark.parametrize("name,c", list(load_cases()))
def test_golden_reference_sampler(name, c):
    t, rnd = split_temps(c["temps"])
    tok, lp, rk, _, _ = ours(c["logits"], t, rnd, c["noise"])
    check_against(tok, lp, rk, c["tokens"], c["logprobs"], c["ranks"], c["logits"], t, c["noise"], rnd)


@pytest.mark.parametrize("rows,vocab,dtype", [(1, 32000, torch.float16), (256, 32000, torch.float16),
                                              (64, 128256, torch.bfloat16), (7, 50272, torch.float32),
                                              (33, 32003, torch.bfloat16), (5, 9, torch.float16)])
def test_against_oracle(rows, vocab, dtype):
    g = torch.Generator().manual_seed(rows * 7 + vocab)
    logits = (torch.randn(rows, vocab, generator=g) * 4).to(dtype)
    noise = torch.empty(rows, vocab).exponential_(generator=g)
    temps = torch.tensor([[1.0, 0.7, 1.3, 1.0][i % 4] for i in range(rows)])
    rnd = torch.tensor([i % 3 != 0 for i in range(rows)])
    t_eff = torch.where(rnd, temps, torch.ones_like(temps))
    want = O.sample_rows(logits, t_eff, rnd, noise)
    tok, lp, rk, probs, lpf = ours(logits, t_eff, rnd, noise, want_probs=True, want_logprobs=True)
    check_against(tok, lp, rk, want[0], want[1], want[2], logits, t_eff.to(dtype).float(), noise, rnd)
    assert torch.allclose(probs.cpu(), want[3], rtol=2e-5, atol=1e-9)
    assert torch.allclose(lpf.cpu(), want[4], rtol=1e-5, atol=2e-5)
    # greedy everywhere, no optional arguments
    tok2, _, rk2, _, _ = ours(logits)
    want2 = O.sample_rows(logits)
    assert torch.equal(tok2.cpu(), want2[0]) and torch.equal(rk2.cpu(), want2[2])
    assert torch.equal(rk2.cpu(), torch.ones_like(rk2.cpu()))


def test_strided_rows_masked_tokens_and_ties():
    g = torch.Generator().manual_seed(3)
    wide = (torch.randn(16, 5000 + 24, generator=g) * 2).to(torch.float16).to(DEV)
    view = wide[:, 8:5008]                      # row stride != vocab, rows 16-byte aligned
    view2 = wide[:, 3:5003]                     # misaligned rows: scalar path
    for v in (view, view2):
        ref = O.sample_rows(v.cpu().contiguous())
        tok, lp, rk, _, _ = ours(v)
        assert torch.equal(tok.cpu(), ref[0]) and torch.equal(rk.cpu(), ref[2])
    # -inf logits are never drawn; equal maxima resolve to the first index (torch.argmax)
    x = torch.full((4, 1024), float("-inf"), dtype=torch.float16)
    x[:, 100] = 1.0
    x[:, 900] = 1.0
    x[:, 17] = 0.5
    noise = torch.empty(4, 1024).exponential_(generator=g)
    tok, lp, rk, _, _ = ours(x, None, torch.ones(4, dtype=torch.bool), noise)
    assert set(tok.cpu().tolist()) <= {17, 100, 900}
    tok, lp, rk, _, _ = ours(x)
    assert tok.cpu().tolist() == [100] * 4 and rk.cpu().tolist() == [1] * 4


def test_in_kernel_noise_distribution_and_determinism():
    # 4096 identical rows of a 16-way distribution, 8 draws each with distinct offsets
    p = torch.tensor([0.3, 0.2, 0.1, 0.1, 0.05, 0.05, 0.05, 0.05, 0.02, 0.02, 0.02, 0.01, 0.01,
                      0.01, 0.005, 0.005], dtype=torch.float64)
    p = p / p.sum()
    logits = torch.log(p).float().repeat(4096, 1)
    rnd = torch.ones(4096, dtype=torch.bool)
    counts = torch.zeros(16, dtype=torch.float64)
    draws = []
    for off in range(8):
        tok, _, _, _, _ = ours(logits, None, rnd, None, seed=1234, offset=off)
        draws.append(tok.cpu())
        counts += torch.bincount(tok.cpu(), minlength=16).double()
    n = counts.sum()
    chi2 = float((((counts - n * p) ** 2) / (n * p)).sum())
    assert chi2 < 50.0, (chi2, counts.tolist())          # 15 dof: p < 1e-5 beyond ~45
    assert not torch.equal(draws[0], draws[1])           # offsets decorrelate steps
    again, _, _, _, _ = ours(logits, None, rnd, None, seed=1234, offset=0)
    assert torch.equal(again.cpu(), draws[0])            # same key, same counter: same draw
    other, _, _, _, _ = ours(logits, None, rnd, None, seed=99, offset=0)
    assert not torch.equal(other.cpu(), draws[0])
    # rows are independent streams: neighbouring rows do not repeat each other
    assert float((draws[0][1:] == draws[0][:-1]).double().mean()) < 0.3
    # temperature sharpens: t = 0.5 squares the odds
    cold, _, _, _, _ = ours(logits, torch.full((4096,), 0.5), rnd, None, seed=7, offset=0)
    q = p ** 2 / (p ** 2).sum()
    c = torch.bincount(cold.cpu(), minlength=16).double()
    assert float((((c - 4096 * q) ** 2) / (4096 * q + 1e-9))[:8].sum()) < 40.0


def test_rejects_cpu_and_bad_layout():
    from vllm_ltr_b200.sampler import sample_rows
    with pytest.raises(RuntimeError):
        sample_rows(torch.zeros(2, 8))
    with pytest.raises(RuntimeError):
        sample_rows(torch.zeros(2, 8, device=DEV).t())


def test_b200_sampler_general_path_on_gpu():
    """Greedy requests with penalties / top-k / top-p / min-p / min_tokens / top-N logprobs through
    B200Sampler with the real kernel, against the reference's own Sampler.forward on the same
    GPU (reference package: oracle/_ref/vllm_py).  Greedy rows are deterministic: tokens, the
    top-N token sets and ranks must match, log-probabilities to fp32 round-off."""
    from tests import sched_replay as R
    if not R.have_reference():
        pytest.skip("reference package not shipped")
    R.import_reference()
    from vllm.model_executor.layers.sampler import Sampler
    from vllm.model_executor.sampling_metadata import SamplingMetadata
    from vllm.sampling_params import SamplingParams, SamplingType
    from vllm.sequence import SequenceData

    from vllm_ltr_b200.sampler import B200Sampler
    vocab = 32000
    specs = [dict(repetition_penalty=1.3, logprobs=3), dict(presence_penalty=0.7, frequency_penalty=0.2),
             dict(top_k=17, logprobs=5), dict(top_p=0.6, logprobs=2), dict(min_p=0.05),
             dict(min_tokens=4, logprobs=1), dict(), dict(top_k=1, top_p=0.5)] * 8
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(len(specs), vocab, generator=g) * 3).to(torch.float16).to(DEV)
    seq_groups, seq_data = [], {}
    cat = {t: [] for t in SamplingType}
    for i, kw in enumerate(specs):
        kw = dict(kw)
        sp = SamplingParams(n=1, temperature=0.0, top_p=kw.pop("top_p", 1.0), max_tokens=16, **kw)
        sp.eos_token_id = 2
        seq_groups.append(([20 + i], sp))
        sd = SequenceData([5, 6, 7, 5, int(logits[i].argmax())])
        for t in (9, int(logits[i].argmax()), 11)[: 1 + i % 3]:
            sd.append_token_id(t, 0.0)
        seq_data[20 + i] = sd
        cat[sp.sampling_type].append((i, len(cat[sp.sampling_type])))
    cat = {t: (torch.tensor(v, dtype=torch.int, device=DEV) if v
               else torch.empty(0, 2, dtype=torch.int, device=DEV)) for t, v in cat.items()}
    md = SamplingMetadata(seq_groups, seq_data, [], torch.arange(len(specs), device=DEV), cat)
    want = Sampler()(logits.clone(), md)
    ours_s = B200Sampler()
    got = ours_s(logits.clone(), md)
    assert ours_s.general_steps == 1 and ours_s.reference_steps == 0
    for a, b in zip(got.outputs, want.outputs):
        sa, sb = a.samples[0], b.samples[0]
        assert sa.output_token == sb.output_token
        assert set(sa.logprobs) == set(sb.logprobs)
        for t in sa.logprobs:
            assert sa.logprobs[t].rank == sb.logprobs[t].rank
            la, lb = sa.logprobs[t].logprob, sb.logprobs[t].logprob
            assert la == lb or abs(la - lb) <= 2e-5 * max(1.0, abs(lb))
