"""Golden vectors for the sampler: runs the REFERENCE's own Sampler.forward
(/root/reference/vllm/model_executor/layers/sampler.py) on CPU tensors and records inputs and
outputs in tests/golden/sampler_ref.npz.  Two interventions, neither touches the arithmetic:
Tensor.exponential_ is replaced while the reference runs so that the noise it draws
(sampler.py:348) is an array we keep, and pinned host staging is turned off (no driver here).

    PYTHONPATH=/root/reference python tests/golden/gen_golden_sampler.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from vllm.model_executor.layers.sampler import Sampler  # noqa: E402
from vllm.model_executor.sampling_metadata import SamplingMetadata  # noqa: E402
from vllm.sampling_params import SamplingParams, SamplingType  # noqa: E402
from vllm.sequence import SequenceData  # noqa: E402

import vllm.model_executor.sampling_metadata as _smd  # noqa: E402

# this container has no CUDA driver: the reference asks for pinned host staging buffers
# (sampling_metadata.py:216), which needs one.  Plain host memory; no arithmetic is affected.
_smd.is_pin_memory_available = lambda: False

HERE = os.path.dirname(os.path.abspath(__file__))


def run_case(name, vocab, dtype, temps, seed):
    """One decode step: len(temps) sequence groups of one sequence; temps[i] == 0 -> greedy."""
    g = torch.Generator().manual_seed(seed)
    R = len(temps)
    logits = (torch.randn(R, vocab, generator=g) * 3.0).to(dtype)
    noise = torch.empty(R, vocab).exponential_(generator=g)
    seq_groups, seq_data = [], {}
    cat = {t: [] for t in SamplingType}
    for i, t in enumerate(temps):
        sp = SamplingParams(n=1, temperature=t, top_p=1.0, ignore_eos=True, max_tokens=16)
        seq_groups.append(([i], sp))
        seq_data[i] = SequenceData([1, 2, 3])
        cat[sp.sampling_type].append((i, len(cat[sp.sampling_type])))
    cat = {t: (torch.tensor(v, dtype=torch.int) if v else torch.empty(0, 2, dtype=torch.int))
           for t, v in cat.items()}
    md = SamplingMetadata(seq_groups=seq_groups, seq_data=seq_data, prompt_lens=[],
                          selected_token_indices=torch.arange(R), categorized_sample_indices=cat)
    rows_random = [i for i, t in enumerate(temps) if t >= 1e-5]
    orig = torch.Tensor.exponential_

    def recorded(self, *a, **k):          # q = torch.empty_like(probs[random rows]); q.exponential_()
        assert self.shape == (len(rows_random), vocab)
        self.copy_(noise[rows_random])
        return self

    torch.Tensor.exponential_ = recorded
    try:
        out = Sampler()(logits.clone(), md)
    finally:
        torch.Tensor.exponential_ = orig
    tokens, lps, ranks = [], [], []
    for grp in out.outputs:
        s = grp.samples[0]
        tokens.append(s.output_token)
        lp = s.logprobs[s.output_token]
        lps.append(lp.logprob)
        ranks.append(lp.rank)
    store = logits.view(torch.int16) if dtype != torch.float32 else logits
    return {f"{name}/logits": store.numpy(), f"{name}/noise": noise.numpy(),
            f"{name}/temps": np.array(temps, dtype=np.float32),
            f"{name}/tokens": np.array(tokens, dtype=np.int64),
            f"{name}/logprobs": np.array(lps, dtype=np.float32),
            f"{name}/ranks": np.array(ranks, dtype=np.int64),
            f"{name}/dtype": np.array(str(dtype).split(".")[-1])}


def main():
    out = {}
    out.update(run_case("f16_mixed", 1000, torch.float16, [0.0, 1.0, 0.7, 1.0, 0.0, 1.3, 1.0, 0.5], 0))
    out.update(run_case("bf16_mixed", 4099, torch.bfloat16, [1.0, 0.0, 1.0, 0.8, 1.0, 1.0], 1))
    out.update(run_case("f32_random", 777, torch.float32, [1.0, 1.0, 0.9, 1.0], 2))
    out.update(run_case("bf16_greedy", 2048, torch.bfloat16, [0.0, 0.0, 0.0], 3))
    np.savez_compressed(os.path.join(HERE, "sampler_ref.npz"), **out)
    print("wrote", os.path.join(HERE, "sampler_ref.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
