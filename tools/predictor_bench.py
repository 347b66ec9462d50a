#!/usr/bin/env python
"""Bulk scoring with the LTR predictor (what the first scheduler tick of synthetic-sharegpt.sh
does: 10 000 queued prompts): tokens/s of OPTPredictor.logits over prompts of the trace, with a
per-kernel breakdown (torch.profiler / CUPTI), plus the single-prompt latency the steady-state
tick pays.  Development tool (run under gpurun).

    python tools/predictor_bench.py [--prompts 2000] [--kind 125m]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prompts", type=int, default=2000)
    ap.add_argument("--kind", default="125m")
    ap.add_argument("--batch-tokens", type=int, default=16384)
    ap.add_argument("--gemm", type=int, default=0, help="1: one-tile-per-CTA GEMM (A/B)")
    ap.add_argument("--no-pdl", action="store_true",
                    help="serialise kernels (per-kernel times in the breakdown are then exact)")
    args = ap.parse_args()
    from vllm_ltr_b200._lib import tune
    from vllm_ltr_b200.predictor import OPTPredictor, dummy_checkpoint
    tune("opt_gemm", args.gemm)
    if args.no_pdl:
        tune("pdl", 0)
    rng = np.random.default_rng(0)
    lens = np.clip(np.round(np.exp(rng.normal(5.0, 1.0, args.prompts))), 4, 2048).astype(int)
    prompts = [rng.integers(3, 50000, int(n)).tolist() for n in lens]
    ntok = int(lens.sum())
    cfg, sd = dummy_checkpoint(args.kind, seed=0)
    pred = OPTPredictor(cfg, sd, device="cuda:0", max_batch_tokens=args.batch_tokens)
    pred.logits(prompts[:200])
    torch.cuda.synchronize()
    for _ in range(2):
        t0 = time.perf_counter()
        pred.logits(prompts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"bulk: {args.prompts} prompts, {ntok} tokens in {dt * 1e3:.1f} ms = "
          f"{ntok / dt / 1e3:.0f}k tokens/s (batches of <= {args.batch_tokens} tokens)")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        pred.logits(prompts)
        torch.cuda.synchronize()
    rows = {}
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type is not None and "cuda" in str(e.device_type).lower():
            rows[e.key] = (t, e.count)
    tot = sum(v[0] for v in rows.values())
    for k, (t, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"  {t / 1e3:9.2f} ms {100 * t / tot:5.1f}%  x{c:<5d} {k[:110]}")
    print(f"  device total {tot / 1e3:.1f} ms")
    for L in (32, 245, 2048):
        p = [prompts[0][:1] * L]
        for _ in range(3):
            pred.logits(p)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            pred.logits(p)
        torch.cuda.synchronize()
        print(f"single prompt of {L} tokens: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per forward")


if __name__ == "__main__":
    main()
