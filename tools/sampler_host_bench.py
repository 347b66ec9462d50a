#!/usr/bin/env python
"""Host time the reference spends per step BEFORE its sampler touches the GPU, for the bench batch
(256 running sequences of the trace, plain temperature sampling as in
benchmark_throughput_original.py:251-259): `SamplingTensors.from_sampling_metadata`
(vllm/model_executor/sampling_metadata.py:79-330) pads every sequence's prompt and output tokens
into [batch, max_len] matrices for penalties that are not requested.  B200Sampler.forward reads
none of that.  CPU only; needs /root/reference.

    python tools/sampler_host_bench.py > profiles/r01_sampler_host.txt
"""
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import sched_replay as R  # noqa: E402

R.import_reference()
logging.disable(logging.CRITICAL)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import vllm.model_executor.sampling_metadata as smd  # noqa: E402
from vllm.model_executor.sampling_metadata import SamplingMetadata, SamplingTensors  # noqa: E402
from vllm.sampling_params import SamplingParams, SamplingType  # noqa: E402
from vllm.sequence import SequenceData  # noqa: E402

smd.is_pin_memory_available = lambda: False      # no CUDA driver in the build container

rng = np.random.default_rng(0)
B = 256
prompt = np.clip(np.rint(np.exp(rng.normal(5.0, 1.0, B))), 4, 2048).astype(int)
out = np.clip(np.rint(np.exp(rng.normal(5.3, 1.1, B))), 4, 8192 - prompt).astype(int)
done = (rng.random(B) * out).astype(int)
seq_groups, seq_data = [], {}
cat = {t: [] for t in SamplingType}
for i in range(B):
    sp = SamplingParams(n=1, temperature=1.0, top_p=1.0, ignore_eos=True, max_tokens=int(out[i]))
    sd = SequenceData(list(range(int(prompt[i]))))
    for _ in range(int(done[i])):
        sd.append_token_id(7, 0.0)
    seq_groups.append(([i], sp))
    seq_data[i] = sd
    cat[sp.sampling_type].append((i, i))
cat = {t: (torch.tensor(v, dtype=torch.int) if v else torch.empty(0, 2, dtype=torch.int))
       for t, v in cat.items()}
md = SamplingMetadata(seq_groups, seq_data, [], torch.arange(B), cat)
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    SamplingTensors.from_sampling_metadata(md, 32000, torch.device("cpu"), torch.float16)
    ts.append(time.perf_counter() - t0)
print(f"reference SamplingTensors.from_sampling_metadata, B={B} sequences of the trace "
      f"(prompt max {int(prompt.max())}, generated max {int(done.max())}): "
      f"median {np.median(ts)*1e3:.1f} ms per step on the host (min {min(ts)*1e3:.1f} ms), before any "
      f"kernel of the sampler runs; B200Sampler.forward builds two [B] lists instead")
