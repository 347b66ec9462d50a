#!/usr/bin/env python
"""reshape_and_cache on the shapes of real engine steps (development tool, run under gpurun).

Cases: a decode step (256 lone tokens), a mixed chunked-prefill step (one 512-token prompt chunk
+ 256 decode tokens - the 768-token budget of the benchmark scripts) and a prefill-only step.
Each op is timed as a CUDA graph of `--n` launches over ROTATING per-layer caches (no launch
finds its destination lines in L2), next to its algorithmic bytes at the measured HBM peak and -
when oracle/_ref/ref_cuda_C is present - the reference kernel on the same inputs.

    python tools/cache_bench.py [--n 32] [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tools"))

DEV = "cuda:0"
MODELS = {"llama-2-7b fp16": (32, torch.float16, "auto"),
          "llama-3-8b bf16": (8, torch.bfloat16, "auto"),
          "llama-3-8b e5m2": (8, torch.bfloat16, "fp8")}


def slots_for(case, nb, bs, rng):
    blocks = rng.permutation(nb)
    if case == "decode 256":
        return [int(b) * bs + int(rng.integers(bs)) for b in blocks[:256]]
    out = []
    n_prefill = 512 if case == "prefill 512 + decode 256" else 768
    start = 5                                   # the chunk continues a prompt mid-block
    for i in range(n_prefill):
        p = start + i
        out.append(int(blocks[p // bs]) * bs + p % bs)
    if case == "prefill 512 + decode 256":
        out += [int(b) * bs + int(rng.integers(bs)) for b in blocks[64:64 + 256]]
    return out


def graph_ms(fn, n, reps=10):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    for _ in range(2):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / (reps * n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from vllm_ltr_b200 import _C
    ref = None
    try:
        from op_sweep import load_ref
        ref = load_ref()
    except Exception as e:                      # noqa: BLE001
        print("reference CUDA build not loaded:", e)
    peak = 6534.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:                           # noqa: BLE001
        pass
    D, bs, nb, hq = 128, 16, 8192, 32
    rows = []
    for mname, (hkv, dt, kv) in MODELS.items():
        es = 1 if kv == "fp8" else 2
        x = 16 // es
        cdt = torch.uint8 if kv == "fp8" else dt
        layers = max(4, min(args.n, (600 << 20) // (2 * nb * hkv * D * bs * es)))
        kcs = [torch.zeros(nb, hkv, D // x, bs, x, dtype=cdt, device=DEV) for _ in range(layers)]
        vcs = [torch.zeros(nb, hkv, D, bs, dtype=cdt, device=DEV) for _ in range(layers)]
        for case in ("decode 256", "prefill 512 + decode 256", "prefill 768"):
            rng = np.random.default_rng(1)
            sl = slots_for(case, nb, bs, rng)
            T = len(sl)
            slots = torch.tensor(sl, dtype=torch.long, device=DEV)
            qkv = torch.randn(T, (hq + 2 * hkv) * D, device=DEV).to(dt)
            key = qkv[:, hq * D:(hq + hkv) * D].view(T, hkv, D)
            val = qkv[:, (hq + hkv) * D:].view(T, hkv, D)
            alg = T * 2 * hkv * D * (2 + es) + 8 * T

            def ours(i):
                _C.cache_ops.reshape_and_cache(key, val, kcs[i % layers], vcs[i % layers], slots,
                                               kv, 1.0)

            ms = graph_ms(ours, args.n)
            row = {"model": mname, "case": case, "tokens": T, "algorithmic_bytes": alg,
                   "ours_us": ms * 1e3, "floor_us": alg / (peak * 1e9) * 1e6}
            if ref is not None:
                def theirs(i):
                    ref.cache_ops.reshape_and_cache(key, val, kcs[i % layers], vcs[i % layers],
                                                    slots, kv, 1.0)
                row["ref_cuda_us"] = graph_ms(theirs, args.n) * 1e3
            rows.append(row)
            print(f"{mname:18s} {case:26s} ours {row['ours_us']:7.2f} us  "
                  f"ref {row.get('ref_cuda_us', float('nan')):7.2f} us  floor {row['floor_us']:5.2f} us "
                  f"({alg / 1e6:.1f} MB)", flush=True)
        del kcs, vcs
        torch.cuda.empty_cache()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
