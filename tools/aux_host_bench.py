#!/usr/bin/env python
"""Host side of scoring a burst of new requests (B200AuxModel.obtain_aux_scores on the 10 000
prompts of the synthetic trace): tokenisation + packing, with the forward replaced by a stand-in
that only packs the token ids the way OPTPredictor._forward does (no GPU needed).  This is the
work the pipelined call overlaps with the forwards on the device; the old path (HF batch call,
per-prompt list copies, generator-fed np.fromiter, all before the first forward) is timed beside
it.  CPU only.

    python tools/aux_host_bench.py > profiles/r02_predictor_host_pipeline.txt
"""
import os
import sys
import tempfile
import time
from itertools import chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import engine_trace as ET  # noqa: E402
from vllm_ltr_b200 import predictor as P  # noqa: E402


class Seq:
    def __init__(self, t):
        self.prompt = t


class Group:
    def __init__(self, t):
        self.seqs_dict = {0: Seq(t)}
        self.aux_model_score = None

    def need_aux_model_score(self):
        return self.aux_model_score is None

    def set_aux_model_score(self, v):
        self.aux_model_score = v


class PackOnly:
    def score_column(self, prompts):
        m = sum(len(p) for p in prompts)
        np.fromiter(chain.from_iterable(prompts), dtype=np.int32, count=m)
        return torch.zeros(len(prompts))


def main():
    from transformers import AutoTokenizer
    d = tempfile.mkdtemp()
    ET._write_tokenizer(d, ET.OPT_VOCAB, 2, 2, prefix_bos=True)
    hf = AutoTokenizer.from_pretrained(d)
    texts = [r[0] for r in ET.make_trace(10000, 32000, 4096)]
    print(f"10000 prompts of the synthetic trace, word-level fast tokenizer, {os.cpu_count()} cores")
    best = None
    for _ in range(3):
        t = time.perf_counter()
        pr = [list(p)[:2048] for p in hf(texts, add_special_tokens=True)["input_ids"]]
        m = sum(map(len, pr))
        np.fromiter((x for p in pr for x in p), dtype=np.int32, count=m)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    print(f"before: hf(texts) + list copies + generator packing, all ahead of the first forward: "
          f"{best:.3f} s ({m} tokens)")
    for chunk in (512, 1024, 2048):
        aux = P.B200AuxModel(PackOnly(), lambda t: hf.encode(t), 2048,
                             P.B200AuxModel._batch_encoder(hf), chunk_prompts=chunk)
        best = None
        for _ in range(3):
            gs = [Group(t) for t in texts]
            t = time.perf_counter()
            aux.obtain_aux_scores(gs)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        print(f"now: chunks of {chunk} prompts, raw encode_batch on the helper thread, packing on "
              f"the caller's: {best:.3f} s of host time in total, overlapped with the forwards "
              f"(first forward enqueued after the first chunk's tokenisation)")


if __name__ == "__main__":
    main()
