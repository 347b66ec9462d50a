#!/usr/bin/env python
"""Scheduler tick cost: reference Scheduler._general_schedule vs the plugin's host logic, pure
Python on one core, at N live requests (baseline 3).  Needs /root/reference.
The CUDA ranker is replaced by a numpy stable argsort here (no GPU in the build container; same
order as Python `sorted`, asserted); its time is reported separately - on the GPU box the rank
kernel call costs ~0.3 ms at N = 10^4 (bench.py tick_ms_alone minus the predictor).

    python tools/sched_tick_bench.py > profiles/r01_scheduler_tick.txt
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
from vllm import SamplingParams  # noqa: E402
from vllm.config import CacheConfig, SchedulerConfig  # noqa: E402
from vllm.core.scheduler import Scheduler  # noqa: E402
from vllm.sequence import Logprob, Sequence, SequenceGroup, SequenceStatus  # noqa: E402

from vllm_ltr_b200 import ltr_scheduler  # noqa: E402

NumpyRanker = R.NumpyRanker


def build(n, plugin):
    sc = SchedulerConfig(768, 256, 8192, enable_chunked_prefill=True, schedule_type="opt-xxx")
    cc = CacheConfig(16, 1.0, 1, "auto")
    cc.num_gpu_blocks, cc.num_cpu_blocks = 40000, 40000
    s = Scheduler(sc, cc, None)
    rng = np.random.default_rng(0)
    score = rng.normal(0, 1, n).astype(np.float16).astype(float)
    s.aux_model = R.StubScorer({str(i): float(score[i]) for i in range(n)})
    if plugin:
        ltr_scheduler.install(s, ranker=NumpyRanker())
    for i in range(n):
        toks = list(range(int(rng.integers(4, 200))))
        s.add_seq_group(SequenceGroup(str(i), [Sequence(i, "", toks, 16)],
                                      SamplingParams(ignore_eos=True, max_tokens=256), 0.0))
    return s


def tick(s):
    metas, o = s.schedule()
    for x in o.scheduled_seq_groups:
        x.seq_group.update_num_computed_tokens(x.token_chunk_size)
        if x.seq_group.get_num_uncomputed_tokens() == 0:
            for seq in x.seq_group.get_seqs(status=SequenceStatus.RUNNING):
                seq.append_token_id(7, {7: Logprob(0.0)})
    return o


print("N live requests | reference ms/tick | plugin host logic ms/tick (of which stand-in ranker) | identical decisions")
for n in (256, 2000, 10000):
    res = {}
    for plugin in (False, True):
        s = build(n, plugin)
        outs = []
        for _ in range(20):
            outs.append([(x.seq_group.request_id, x.token_chunk_size) for x in tick(s).scheduled_seq_groups])
        ts = []
        NumpyRanker.seconds = 0.0
        for _ in range(8):
            t0 = time.perf_counter()
            o = s._schedule()     # the tick proper (metadata building in schedule() is outside the path)
            ts.append(time.perf_counter() - t0)
            # undo nothing: keep going with the engine bookkeeping
            for x in o.scheduled_seq_groups:
                x.seq_group.update_num_computed_tokens(x.token_chunk_size)
                if x.seq_group.get_num_uncomputed_tokens() == 0:
                    for seq in x.seq_group.get_seqs(status=SequenceStatus.RUNNING):
                        seq.append_token_id(7, {7: Logprob(0.0)})
            outs.append([(x.seq_group.request_id, x.token_chunk_size) for x in o.scheduled_seq_groups])
        res[plugin] = (float(np.median(ts)) * 1e3, outs, NumpyRanker.seconds / 8 * 1e3)
    print(f"{n:>15d} | {res[False][0]:>17.2f} | {res[True][0]:>25.2f} ({res[True][2]:.2f}) | "
          f"{res[False][1] == res[True][1]}")
