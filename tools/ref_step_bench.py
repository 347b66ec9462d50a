#!/usr/bin/env python
"""The decode hot-path step with OUR kernels next to the same step with the REFERENCE's CUDA
kernels (csrc/*.cu recompiled unmodified for sm_100a into oracle/_ref/ref_cuda_C by
oracle/Makefile) - the "reference CUDA build" target of BASELINE.md section 3.1 at step level.
Both go through tools/hotpath.py DecodeHotPath (the call sequence of llama.py:215-238) with
the same shapes, pools and inputs; only the extension module behind `vllm._C` differs.
Development tool (run under gpurun), like tools/op_sweep.py.

    python tools/ref_step_bench.py [--models llama-2-7b,llama-3-8b,llama-3-8b-fp8kv] [--batch 256]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench import build_inputs, trace_contexts  # noqa: E402
from op_sweep import load_ref  # noqa: E402

DEV = "cuda:0"


def time_steps(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="llama-2-7b,llama-3-8b,llama-3-8b-fp8kv")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from vllm_ltr_b200 import _C
    from hotpath import SHAPES, DecodeHotPath, ModelShape
    ref = load_ref()
    assert ref is not None, "oracle/_ref/ref_cuda_C not built (make -C oracle ref_cuda)"
    rows = []
    for name in args.models.split(","):
        s = ModelShape(**{**SHAPES[name].__dict__})
        s.dtype = torch.bfloat16
        ctx = trace_contexts(args.batch)
        nb = int(((ctx + 15) // 16).sum()) + 64
        inp = {k: v.to(DEV) for k, v in build_inputs(ctx, 16, nb, 0).items()}
        h = (torch.randn(args.batch, s.hidden, device=DEV) * 0.5).to(s.dtype)
        res = {}
        for label, mod in (("ours", _C), ("reference CUDA kernels", ref)):
            hp = DecodeHotPath(s, nb, args.batch, DEV, C=mod)
            a = (h, inp["positions"], inp["slot_mapping"], inp["block_tables"], inp["context_lens"],
                 int(ctx.max()))
            eager = time_steps(lambda: hp.step(*a), args.steps)
            g, _ = hp.capture(*a)
            graph = time_steps(g.replay, args.steps)
            res[label] = (eager, graph)
            del hp, g
            torch.cuda.empty_cache()
        o, r = res["ours"], res["reference CUDA kernels"]
        print(f"{name} B={args.batch}: ours {o[0]:.3f} ms eager / {o[1]:.3f} ms graph "
              f"({args.batch/o[1]*1e3:.0f} tokens/s) | reference kernels {r[0]:.3f} ms eager / "
              f"{r[1]:.3f} ms graph ({args.batch/r[1]*1e3:.0f} tokens/s) | x{r[1]/o[1]:.2f} (graph), "
              f"x{r[0]/o[0]:.2f} (eager)", flush=True)
        rows.append(dict(model=name, batch=args.batch, ours_eager_ms=o[0], ours_graph_ms=o[1],
                         ref_eager_ms=r[0], ref_graph_ms=r[1]))
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
