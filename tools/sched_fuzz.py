#!/usr/bin/env python
"""Differential fuzz of the LTR tick: random populations, budgets, KV pool sizes, starvation settings
and aborts between ticks, each replayed through the reference ``Scheduler`` and through the plugin
(tests/sched_replay.py); every tick's record (scheduled requests, block-table sizes, metadata
fields, swap maps, queue orders, idle / runs / pri of every live request) must be identical, and
a configuration the reference rejects must be rejected the same way.  Needs /root/reference.

    python tools/sched_fuzz.py [first_seed [count]]
"""
import os
import random
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import logging; logging.disable(logging.CRITICAL)
from tests import sched_replay as R
bad = 0
first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for seed in range(first, first + count):
    rnd = random.Random(seed)
    n = rnd.choice([30, 80, 150, 250])
    stype = rnd.choice(["opt-xxx", "opt-xxx", "opt-xxx", "opt-xxx-starv5-period3", "opt-xxx-starv12-period2"])
    gpu = rnd.choice([64, 120, 200, 400, 4096])
    cpu = rnd.choice([4096, 4096, 300])
    mt = rnd.choice([32, 64, 128, 768]); ms = rnd.choice([4, 8, 16, 64, 256])
    aborts = {}
    for _ in range(rnd.randint(0, 8)):
        aborts.setdefault(rnd.randint(0, 60), []).extend(rnd.sample(range(n + 5), rnd.randint(1, 6)))
    kw = dict(max_tokens=mt, max_seqs=ms, aborts=aborts, max_ticks=600)
    try:
        ref = R.run(n, seed, stype, gpu, cpu, use_plugin=False, **kw); e1 = None
    except BaseException as e:
        ref, e1 = None, (type(e).__name__, str(e)[:80])
    try:
        ours = R.run(n, seed, stype, gpu, cpu, use_plugin=True, ranker=R.NumpyRanker(), **kw); e2 = None
    except BaseException as e:
        ours, e2 = None, (type(e).__name__, str(e)[:80])
    ok = (ref == ours) and (e1 == e2 or (e1 and e2 and e1[0] == e2[0]))
    if not ok: bad += 1
    print(seed, n, stype, gpu, cpu, mt, ms, len(aborts), "ticks", None if ref is None else len(ref), "err", e1, e2, "OK" if ok else "MISMATCH", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
