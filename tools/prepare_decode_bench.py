#!/usr/bin/env python
"""Host time of ModelRunner._prepare_decode (vllm/worker/model_runner.py:424-548) vs
vllm_ltr_b200.decode_inputs.prepare_decode for the bench batch (256 running sequences of the
trace).  CPU only; needs /root/reference.

    python tools/prepare_decode_bench.py > profiles/r01_prepare_decode.txt
"""
import logging
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import sched_replay as R  # noqa: E402

R.import_reference()
logging.disable(logging.CRITICAL)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from vllm import SamplingParams  # noqa: E402
from vllm.sequence import SequenceData, SequenceGroupMetadata  # noqa: E402
from vllm.worker.model_runner import ModelRunner  # noqa: E402

from bench import trace_contexts  # noqa: E402
from vllm_ltr_b200 import decode_inputs as ours  # noqa: E402

for B in (64, 256):
    ctx = trace_contexts(B)
    metas, nb = [], 0
    for i, c in enumerate(ctx):
        n = int(c)
        sd = SequenceData(list(range(100, 100 + min(n, 50))))
        for _ in range(n - min(n, 50)):
            sd.append_token_id(7, 0.0)
        nblk = (n + 15) // 16
        metas.append(SequenceGroupMetadata(str(i), False, {i: sd}, SamplingParams(),
                                           {i: list(range(nb, nb + nblk))}, token_chunk_size=1))
        nb += nblk
    fake = types.SimpleNamespace(
        sliding_window=None, block_size=16, device="cpu",
        model_config=types.SimpleNamespace(enforce_eager=True), max_context_len_to_capture=8192,
        attn_backend=types.SimpleNamespace(make_metadata=lambda **kw: types.SimpleNamespace(**kw)))
    res = {}
    for name, fn in (("reference", ModelRunner._prepare_decode), ("ours", ours.prepare_decode)):
        out = fn(fake, metas)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            fn(fake, metas)
            ts.append(time.perf_counter() - t0)
        res[name] = (float(np.median(ts)) * 1e3, out)
    same = (res["ours"][1].slot_mapping == res["reference"][1].slot_mapping and torch.equal(
        res["ours"][1].attn_metadata.block_tables, res["reference"][1].attn_metadata.block_tables))
    print(f"_prepare_decode, B={B} sequences of the trace ({int(ctx.sum())} context tokens): reference "
          f"{res['reference'][0]:.2f} ms, ours {res['ours'][0]:.2f} ms "
          f"(x{res['reference'][0]/res['ours'][0]:.1f}), identical output: {same}")
