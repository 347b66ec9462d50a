#!/usr/bin/env python
"""Paged-attention launch-policy sweep and per-warp timeline (development tool, run under gpurun).

Timing here is the in-step situation of bench.py, not the cold single launch of op_sweep.py: the
kernel is launched back to back over ROTATING KV pools whose total size exceeds L2 (so no launch
finds its KV in cache), 24 launches between one pair of CUDA events, no L2 flush in between.

  python tools/attn_tune.py [--cases 8b-gqa-bf16:256:trace,...] [--grid default|wide] [--trace]

--trace prints, for the default policy, where the time of one launch goes, from globaltimer
stamps every warp leaves (vltr_tune("attn_trace", ptr)): launch -> plan built -> first KV tile
landed -> warp finished.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import trace_contexts  # noqa: E402

DEV = "cuda:0"
CFGS = {"7b-mha-bf16": (32, 32, torch.bfloat16, "auto"),
        "8b-gqa-bf16": (32, 8, torch.bfloat16, "auto"),
        "8b-gqa-fp8": (32, 8, torch.bfloat16, "fp8"),
        "70b-tp8-bf16": (8, 1, torch.bfloat16, "auto"),
        "7b-tp8-bf16": (4, 4, torch.bfloat16, "auto"),
        "7b-tp2-bf16": (16, 16, torch.bfloat16, "auto"),
        "8b-tp8-bf16": (4, 1, torch.bfloat16, "auto")}


def make_case(name, B, mode, seed=0, min_total_bytes=400 << 20, max_pools=12):
    hq, hkv, dt, kv = CFGS[name]
    D, bs = 128, 16
    ctx = trace_contexts(B, seed) if mode == "trace" else np.full(B, int(mode), dtype=np.int64)
    nblk = (ctx + bs - 1) // bs
    total = int(nblk.sum())
    nb = total + 16
    es = 1 if kv == "fp8" else 2
    x = 16 // es
    pool_bytes = 2 * nb * hkv * D * bs * es
    npools = int(min(max_pools, max(2, -(-min_total_bytes // pool_bytes))))
    g = torch.Generator(device=DEV).manual_seed(seed)
    pools = []
    for _ in range(npools):
        if kv == "fp8":
            kc = torch.randint(0, 0x38, (nb, hkv, D // x, bs, x), generator=g, device=DEV, dtype=torch.uint8)
            vc = torch.randint(0, 0x38, (nb, hkv, D, bs), generator=g, device=DEV, dtype=torch.uint8)
        else:
            kc = ((torch.rand(nb, hkv, D // x, bs, x, generator=g, device=DEV) - 0.5) * 0.2).to(dt)
            vc = ((torch.rand(nb, hkv, D, bs, generator=g, device=DEV) - 0.5) * 0.2).to(dt)
        pools.append((kc, vc))
    perm = np.random.default_rng(seed).permutation(nb)[:total]
    bt = np.zeros((B, int(nblk.max())), dtype=np.int32)
    o = 0
    for i in range(B):
        bt[i, :nblk[i]] = perm[o:o + nblk[i]]
        o += nblk[i]
    q = (torch.randn(B, hq, D, generator=g, device=DEV) * 0.1).to(dt)
    byts = int(ctx.sum()) * 2 * hkv * D * es + 2 * B * hq * D * 2 + 4 * total + 4 * B
    return dict(q=q, out=torch.empty_like(q), pools=pools, bt=torch.from_numpy(bt).to(DEV),
                cl=torch.from_numpy(ctx.astype(np.int32)).to(DEV), mx=int(ctx.max()), hkv=hkv,
                kv=kv, bytes=byts, npools=npools)


def launch(ops, c, i):
    kc, vc = c["pools"][i % c["npools"]]
    ops.paged_attention_v1(c["out"], c["q"], kc, vc, c["hkv"], 128 ** -0.5, c["bt"], c["cl"], 16,
                           c["mx"], None, c["kv"], 1.0)


def time_case(ops, c, iters=24, reps=3):
    for i in range(4):
        launch(ops, c, i)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            launch(ops, c, i)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best      # ms per launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="7b-mha-bf16:256:trace,8b-gqa-bf16:256:trace,8b-gqa-fp8:256:trace,"
                                       "70b-tp8-bf16:256:trace,8b-gqa-fp8:256:2048,8b-gqa-bf16:64:trace")
    ap.add_argument("--grid", default="default")
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--ref", action="store_true", help="also time the reference kernel (oracle/_ref)")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from vllm_ltr_b200 import _C
    from vllm_ltr_b200._lib import tune
    ref = None
    if args.ref:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from op_sweep import load_ref
        ref = load_ref()
    peak = 6534.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    grids = {
        "default": [{}],
        "wide": [{}, {"pdl": 0}, {"attn_ctas": 3}, {"attn_ctas": 4}, {"attn_ctas": 4, "attn_warps": 2},
                 {"attn_stages": 4}, {"attn_stages": 3}, {"attn_slots": 4}, {"attn_slots": 16},
                 {"attn_slots": 32}, {"attn_cap_tokens": 256}, {"attn_cap_tokens": 128},
                 {"attn_cap_tokens": 256, "attn_slots": 16}, {"attn_ctas": 3, "attn_slots": 12},
                 {"attn_ctas": 4, "attn_slots": 16}, {"attn_ctas": 4, "attn_slots": 16, "attn_cap_tokens": 256}],
        "r3": [{}, {"attn_slots": 12}, {"attn_slots": 16}, {"attn_cap_tokens": 256},
               {"attn_min_tokens": 128}, {"attn_min_tokens": 256}],
        "ctas": [{}, {"attn_ctas": 2}, {"attn_ctas": 2, "attn_stages": 3},
                 {"attn_ctas": 2, "attn_stages": 4}, {"attn_ctas": 3, "attn_stages": 3}],
        "cap": [{}, {"attn_cap_tokens": 384}, {"attn_cap_tokens": 256}, {"attn_cap_tokens": 128},
                {"attn_ctas": 2, "attn_cap_tokens": 256}],
        "fp8": [{}, {"attn_stages": 3}, {"attn_stages": 4}, {"attn_ctas": 3}],
        # round 2: the two-launch (split) mode of the latency-bound regime
        "split": [{}, {"attn_split": -1}, {"attn_split": 1}, {"attn_split": 1, "attn_min_tokens": 32},
                  {"attn_split": 1, "attn_min_tokens": 128}, {"attn_split": 1, "attn_min_tokens": 256},
                  {"attn_split": 1, "attn_ctas": 2}, {"attn_split": 1, "attn_ctas": 4}],
    }[args.grid]
    keys = ("pdl", "attn_ctas", "attn_warps", "attn_stages", "attn_slots", "attn_cap_tokens",
            "attn_min_tokens", "attn_split")
    rows = []
    for spec in args.cases.split(","):
        name, B, mode = spec.split(":")
        c = make_case(name, int(B), mode)
        print(f"== {spec}: {c['bytes']/1e6:.1f} MB algorithmic, {c['npools']} rotating pools", flush=True)
        if ref is not None:
            ms = time_case(ref.ops, c)
            print(f"   reference kernel (sm_100a build)        {ms*1e3:8.1f} us  {c['bytes']/ms/1e6:7.0f} GB/s", flush=True)
            rows.append(dict(case=spec, knobs="reference", us=ms * 1e3))
        for knobs in grids:
            for k in keys:
                tune(k, 1 if k == "pdl" else 0)
            for k, v in knobs.items():
                tune(k, v)
            try:
                ms = time_case(_C.ops, c)
            except RuntimeError as e:
                print(f"   {knobs}: {e}")
                continue
            gbs = c["bytes"] / ms / 1e6
            print(f"   {str(knobs) if knobs else 'library default':40s} {ms*1e3:8.1f} us  {gbs:7.0f} GB/s  "
                  f"({gbs/peak*100:5.1f}% of {peak:.0f})", flush=True)
            rows.append(dict(case=spec, knobs=knobs, us=ms * 1e3, gbs=gbs, frac=gbs / peak))
        for k in keys:
            tune(k, 1 if k == "pdl" else 0)
        if args.trace:
            nw = 148 * 16 * 2
            buf = torch.zeros(nw * 16, dtype=torch.int64, device=DEV)
            for i in range(3):
                launch(_C.ops, c, i)
            torch.cuda.synchronize()
            tune("attn_trace", buf.data_ptr())
            for i in range(4):                 # back to back: the buffer keeps the last launch
                launch(_C.ops, c, i)
            torch.cuda.synchronize()
            tune("attn_trace", 0)
            t = buf.cpu().numpy().reshape(nw, 16)
            t = t[t[:, 0] != 0]
            if len(t) == 0:
                print("   (no timeline: build the library with `make -C vllm_ltr_b200/csrc clean all "
                      "EXTRA=-DVLTR_ATTN_TRACE=1`)")
                continue
            t0 = t[:, 0].min()
            start, plan, first, end = (t[:, i] - t0 for i in range(4))
            worked = t[:, 5] > 0
            pct = lambda a: "min %6.1f  p50 %6.1f  p90 %6.1f  max %6.1f" % tuple(
                np.percentile(a, [0, 50, 90, 100]) / 1e3)
            print(f"   timeline of one launch (us after the first warp started; {len(t)} warps, "
                  f"{int(worked.sum())} got work):")
            print(f"     warp start            {pct(start)}")
            print(f"     plan built, deps met  {pct(plan)}")
            if worked.any():
                print(f"     first KV tile landed  {pct(first[worked])}")
            print(f"     warp finished         {pct(end)}")
            w = worked & (t[:, 7] > 0)
            if w.any():
                frac = t[w, 6] / t[w, 7]
                cyc = (t[w, 7] - t[w, 6]) / t[w, 5]
                print(f"     tile loop: waiting for KV {np.median(frac)*100:.0f}% of the time (p10 "
                      f"{np.percentile(frac,10)*100:.0f}%, p90 {np.percentile(frac,90)*100:.0f}%); "
                      f"compute {np.median(cyc):.0f} SM cycles per tile")
            if worked.any():
                tw = t[worked]
                for label, sel in (("single-partition", tw[:, 12] == 1), ("multi-partition", tw[:, 12] > 1)):
                    if sel.any():
                        x = tw[sel]
                        print(f"     last item of a warp, {label} ({int(sel.sum())} warps, {np.median(x[:,13]):.0f} tiles p50): "
                              f"tiles {np.median(x[:,9]-x[:,8])/1e3:.1f} us (max {np.max(x[:,9]-x[:,8])/1e3:.1f}), "
                              f"store {np.median(x[:,10]-x[:,9])/1e3:.1f} us, "
                              f"ticket+merge {np.median(x[:,11]-x[:,10])/1e3:.1f} us (max {np.max(x[:,11]-x[:,10])/1e3:.1f}), "
                              f"exit after {np.median(x[:,3]-x[:,11])/1e3:.1f} us")
            print(f"     items per warp  min {t[:,4].min()} p50 {np.median(t[:,4]):.0f} max {t[:,4].max()};"
                  f"  KV tiles per warp  min {t[:,5].min()} p50 {np.median(t[:,5]):.0f} max {t[:,5].max()}",
                  flush=True)
            rows.append(dict(case=spec, trace=dict(
                start=np.percentile(start, [0, 50, 90, 100]).tolist(),
                plan=np.percentile(plan, [0, 50, 90, 100]).tolist(),
                first=np.percentile(first[worked], [0, 50, 90, 100]).tolist() if worked.any() else None,
                end=np.percentile(end, [0, 50, 90, 100]).tolist())))
        del c
        torch.cuda.empty_cache()
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1, default=str)


if __name__ == "__main__":
    main()
