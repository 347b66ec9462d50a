"""CPU: the sampler oracle against the reference's own Sampler.forward (golden vectors made by
tests/golden/gen_golden_sampler.py) and, where /root/reference is present, the host logic of
B200Sampler.forward against the reference Sampler object by object."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sampler_ref.npz")
_DT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def load_cases():
    z = np.load(GOLD)
    for name in sorted({k.split("/")[0] for k in z.files}):
        dt = _DT[str(z[f"{name}/dtype"])]
        raw = torch.from_numpy(z[f"{name}/logits"])
        logits = raw if dt == torch.float32 else raw.view(dt)
        yield name, dict(logits=logits, noise=torch.from_numpy(z[f"{name}/noise"]),
                         temps=torch.from_numpy(z[f"{name}/temps"]),
                         tokens=torch.from_numpy(z[f"{name}/tokens"]),
                         logprobs=torch.from_numpy(z[f"{name}/logprobs"]),
                         ranks=torch.from_numpy(z[f"{name}/ranks"]))


def split_temps(temps):
    """the reference maps temperature < 1e-5 to greedy with temperature 1 (sampling_metadata.py:133)"""
    rnd = temps >= 1e-5
    return torch.where(rnd, temps, torch.ones_like(temps)), rnd


@pytest.mark.parametrize("name,c", list(load_cases()))
def test_oracle_matches_reference_sampler(name, c):
    t, rnd = split_temps(c["temps"])
    tok, lp, rk, _, _ = O.sample_rows(c["logits"], t, rnd, c["noise"])
    assert torch.equal(tok, c["tokens"])
    assert torch.equal(rk, c["ranks"])
    assert torch.equal(lp, c["logprobs"])          # same torch ops on the same machine: exact


@pytest.mark.skipif(not os.path.isdir("/root/reference/vllm"), reason="needs the reference checkout")
def test_b200_sampler_host_logic_against_reference(monkeypatch):
    sys.path.insert(0, "/root/reference")
    import vllm.model_executor.sampling_metadata as smd
    from vllm.model_executor.layers.sampler import Sampler
    from vllm.model_executor.sampling_metadata import SamplingMetadata
    from vllm.sampling_params import SamplingParams, SamplingType
    from vllm.sequence import SequenceData
    monkeypatch.setattr(smd, "is_pin_memory_available", lambda: False)
    sampler_mod = pytest.importorskip("vllm_ltr_b200.sampler")

    vocab, temps = 501, [1.0, 0.0, 0.7, 1.0, 0.0]
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(len(temps), vocab, generator=g) * 3).to(torch.float16)
    noise = torch.empty(len(temps), vocab).exponential_(generator=g)
    seq_groups, seq_data = [], {}
    cat = {t: [] for t in SamplingType}
    for i, t in enumerate(temps):
        sp = SamplingParams(n=1, temperature=t, top_p=1.0, ignore_eos=True, max_tokens=8)
        seq_groups.append(([10 + i], sp))
        seq_data[10 + i] = SequenceData([1, 2, 3])
        cat[sp.sampling_type].append((i, len(cat[sp.sampling_type])))
    cat = {t: (torch.tensor(v, dtype=torch.int) if v else torch.empty(0, 2, dtype=torch.int))
           for t, v in cat.items()}
    md = SamplingMetadata(seq_groups, seq_data, [], torch.arange(len(temps)), cat)
    rows_random = [i for i, t in enumerate(temps) if t >= 1e-5]
    monkeypatch.setattr(torch.Tensor, "exponential_",
                        lambda self, *a, **k: self.copy_(noise[rows_random]))
    want = Sampler()(logits.clone(), md, pred_scores=torch.arange(len(temps), dtype=torch.float))

    # the kernel call replaced by the oracle: this test is about the host side of forward()
    def fake_rows(lg, temperatures=None, random=None, noise_=None, seed=0, offset=0, **kw):
        tok, lp, rk, _, _ = O.sample_rows(lg, temperatures, random, noise)
        return tok, lp, rk, None, None
    monkeypatch.setattr(sampler_mod, "sample_rows", fake_rows)
    got = sampler_mod.B200Sampler()(logits.clone(), md, pred_scores=torch.arange(len(temps), dtype=torch.float))
    assert len(got.outputs) == len(want.outputs)
    for a, b in zip(got.outputs, want.outputs):
        assert a.prompt_logprobs == b.prompt_logprobs and a.pred_score == b.pred_score
        assert len(a.samples) == len(b.samples) == 1
        sa, sb = a.samples[0], b.samples[0]
        assert (sa.parent_seq_id, sa.output_token) == (sb.parent_seq_id, sb.output_token)
        assert set(sa.logprobs) == set(sb.logprobs)
        la, lb = sa.logprobs[sa.output_token], sb.logprobs[sb.output_token]
        assert la.rank == lb.rank and abs(la.logprob - lb.logprob) <= 1e-6

    bad = SamplingParams(n=1, temperature=1.0, use_beam_search=False, best_of=1, seed=7, max_tokens=8)
    assert bad.sampling_type == SamplingType.RANDOM_SEED
    md_bad = SamplingMetadata([([0], bad)], {0: SequenceData([1])}, [], torch.arange(1), cat)
    with pytest.raises(NotImplementedError, match="sampling_type"):
        sampler_mod.B200Sampler()(logits[:1].clone(), md_bad)
    # with the engine's own Sampler behind it, such a step is handed over whole
    calls = []

    class Ref(torch.nn.Module):
        def forward(self, lg, smd_, pred_scores=None, aux_model_scores=None):
            calls.append(lg.shape)
            return "REFERENCE OUTPUT"

    s2 = sampler_mod.B200Sampler(reference=Ref())
    assert s2(logits[:1].clone(), md_bad) == "REFERENCE OUTPUT"
    assert s2.reference_steps == 1 and calls == [logits[:1].shape]


@pytest.mark.skipif(not os.path.isdir("/root/reference/vllm"), reason="needs the reference checkout")
def test_b200_sampler_general_path_against_reference(monkeypatch):
    """Penalties, top-k / top-p / min-p, min_tokens and top-N logprobs: B200Sampler applies the
    reference's own logits edits and then the fused kernel (the oracle stands in for it here); the
    SamplerOutput must equal the reference Sampler's, object by object, for greedy rows and for
    random rows drawn from the same recorded exponential noise."""
    sys.path.insert(0, "/root/reference")
    import vllm.model_executor.sampling_metadata as smd
    from vllm.model_executor.layers.sampler import Sampler
    from vllm.model_executor.sampling_metadata import SamplingMetadata
    from vllm.sampling_params import SamplingParams, SamplingType
    from vllm.sequence import SequenceData
    monkeypatch.setattr(smd, "is_pin_memory_available", lambda: False)
    sampler_mod = pytest.importorskip("vllm_ltr_b200.sampler")
    vocab = 301
    specs = [dict(temperature=0.0, repetition_penalty=1.3, logprobs=3),
             dict(temperature=0.0, presence_penalty=0.7, frequency_penalty=0.2),
             dict(temperature=0.8, top_k=17), dict(temperature=1.0, top_p=0.6, logprobs=2),
             dict(temperature=0.9, min_p=0.05), dict(temperature=0.0, min_tokens=4, logprobs=1),
             dict(temperature=1.0), dict(temperature=0.0)]
    g = torch.Generator().manual_seed(9)
    logits = (torch.randn(len(specs), vocab, generator=g) * 3).to(torch.float32)
    noise = torch.empty(len(specs), vocab).exponential_(generator=g)
    seq_groups, seq_data = [], {}
    cat = {t: [] for t in SamplingType}
    for i, kw in enumerate(specs):
        sp = SamplingParams(n=1, top_p=kw.pop("top_p", 1.0), ignore_eos=False, max_tokens=16, **kw)
        sp.eos_token_id = 2
        seq_groups.append(([20 + i], sp))
        sd = SequenceData([5, 6, 7, 5])
        for t in (9, 5, 11)[: 1 + i % 3]:
            sd.append_token_id(t, 0.0)
        seq_data[20 + i] = sd
        cat[sp.sampling_type].append((i, len(cat[sp.sampling_type])))
    cat = {t: (torch.tensor(v, dtype=torch.int) if v else torch.empty(0, 2, dtype=torch.int))
           for t, v in cat.items()}
    md = SamplingMetadata(seq_groups, seq_data, [], torch.arange(len(specs)), cat)
    rows_random = [i for i, (_, sp) in enumerate(seq_groups) if sp.sampling_type == SamplingType.RANDOM]
    monkeypatch.setattr(torch.Tensor, "exponential_",
                        lambda self, *a, **k: self.copy_(noise[rows_random]))
    want = Sampler()(logits.clone(), md)

    def fake_rows(lg, temperatures=None, random=None, noise_=None, seed=0, offset=0,
                  want_probs=False, want_logprobs=False):
        tok, lp, rk, _, lpf = O.sample_rows(lg, temperatures, random, noise)
        return tok, lp, rk, None, (lpf if want_logprobs else None)
    monkeypatch.setattr(sampler_mod, "sample_rows", fake_rows)
    ours = sampler_mod.B200Sampler()
    got = ours(logits.clone(), md)
    assert ours.general_steps == 1 and ours.reference_steps == 0
    assert len(got.outputs) == len(want.outputs) == len(specs)
    for a, b in zip(got.outputs, want.outputs):
        sa, sb = a.samples[0], b.samples[0]
        assert (sa.parent_seq_id, sa.output_token) == (sb.parent_seq_id, sb.output_token)
        assert set(sa.logprobs) == set(sb.logprobs)
        for t in sa.logprobs:
            assert sa.logprobs[t].rank == sb.logprobs[t].rank
            assert abs(sa.logprobs[t].logprob - sb.logprobs[t].logprob) <= 1e-6
