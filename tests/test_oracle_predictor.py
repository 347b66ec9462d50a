"""Pins oracle/opt_oracle.py against HF transformers' OPTForSequenceClassification (the class the
reference's trainer saves its predictors from, train/trainer.py:214-226) on seeded random weights
in the OPT-125M (pre-LN) and OPT-350M (post-LN, project_in/out) layouts.  CPU only."""
import pytest
import torch

from oracle import opt_oracle as O


@pytest.mark.parametrize("kind,layers", [("125m", 3), ("350m", 2)])
@pytest.mark.parametrize("num_labels", [1, 5])
def test_oracle_matches_hf(kind, layers, num_labels):
    transformers = pytest.importorskip("transformers")
    cfg, sd = O.random_opt(kind, seed=3, num_labels=num_labels, vocab=1000, max_pos=128,
                           layers=layers)
    hf_cfg = transformers.OPTConfig(**{k: v for k, v in cfg.items()})
    model = transformers.OPTForSequenceClassification(hf_cfg).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("lm_head" in m for m in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(0)
    prompts = [torch.randint(3, 1000, (n,), generator=g).tolist() for n in (1, 2, 17, 64, 100)]
    ours = O.opt_logits(cfg, sd, prompts)
    with torch.no_grad():
        theirs = torch.cat([model(input_ids=torch.tensor([p])).logits for p in prompts])
    assert torch.allclose(ours, theirs.float(), atol=2e-5, rtol=1e-4), (ours - theirs).abs().max()


def test_scores_are_model_dtype_floats():
    cfg, sd = O.random_opt("125m", seed=1, vocab=500, max_pos=64, layers=1)
    s = O.opt_scores(cfg, sd, [[5, 6, 7], [9]])
    assert all(float(torch.tensor(x, dtype=torch.float16)) == x for x in s)
    cfg5, sd5 = O.random_opt("125m", seed=1, num_labels=4, vocab=500, max_pos=64, layers=1)
    assert all(x in (0.0, 1.0, 2.0, 3.0) for x in O.opt_scores(cfg5, sd5, [[5, 6, 7], [9]]))
