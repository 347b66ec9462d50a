#!/usr/bin/env python
"""A/B of the plugin's tick host logic between a git revision and the working tree, alternating
tick by tick in ONE process on the same request population (the build container's cores are
noisy: medians move by 2x between runs, the A/B order within a run does not).  The stand-in
ranker's time (numpy argsort; the product's rank kernel runs on the GPU) is subtracted from both.
Needs /root/reference.

    python tools/sched_tick_ab.py <git-rev> >> profiles/r02_scheduler_tick_v2.txt
"""
import os
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
src = open(os.path.join(ROOT, "tools", "sched_tick_bench.py")).read().split('print("N live requests')[0]
g = {"__file__": os.path.join(ROOT, "tools", "sched_tick_bench.py"), "__name__": "x"}
exec(compile(src, "stb", "exec"), g)
import numpy as np
from vllm.sequence import Logprob, SequenceStatus
import vllm_ltr_b200.ltr_scheduler as NEW
rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
old_src = subprocess.check_output(["git", "-C", ROOT, "show", f"{rev}:vllm_ltr_b200/ltr_scheduler.py"]).decode()
OLD = types.ModuleType("vllm_ltr_b200.ltr_scheduler_old"); OLD.__package__ = "vllm_ltr_b200"
exec(compile(old_src, "old_ltr", "exec"), OLD.__dict__)
R = g["R"]
def build(n, mod):
    L = g["ltr_scheduler"]
    g["ltr_scheduler"] = mod
    try: return g["build"](n, True)
    finally: g["ltr_scheduler"] = L
def adv(o):
    for x in o.scheduled_seq_groups:
        x.seq_group.update_num_computed_tokens(x.token_chunk_size)
        if x.seq_group.get_num_uncomputed_tokens() == 0:
            for seq in x.seq_group.get_seqs(status=SequenceStatus.RUNNING):
                seq.append_token_id(7, {7: Logprob(0.0)})
print(f"tick host logic, revision {rev} (old) vs working tree (new), 30 alternating ticks each")
for n in (256, 2000, 10000):
    so, sn = build(n, OLD), build(n, NEW)
    for _ in range(20): adv(so._schedule()); adv(sn._schedule())
    to, tn = [], []
    for _ in range(30):
        R.NumpyRanker.seconds = 0
        t=time.perf_counter(); o1=so._schedule(); to.append(time.perf_counter()-t - R.NumpyRanker.seconds); adv(o1)
        R.NumpyRanker.seconds = 0
        t=time.perf_counter(); o2=sn._schedule(); tn.append(time.perf_counter()-t - R.NumpyRanker.seconds); adv(o2)
        assert [(x.seq_group.request_id, x.token_chunk_size) for x in o1.scheduled_seq_groups] == [(x.seq_group.request_id, x.token_chunk_size) for x in o2.scheduled_seq_groups]
    print(f"N={n:6d}  old {np.median(to)*1e3:6.2f} ms  new {np.median(tn)*1e3:6.2f} ms   (host logic without the stand-in ranker; min old {min(to)*1e3:.2f} new {min(tn)*1e3:.2f})")
