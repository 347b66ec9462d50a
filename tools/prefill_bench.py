#!/usr/bin/env python
"""Prefill attention under chunked prefill: vltr_context_attention_fwd (csrc/prefill_attention.cu)
next to the reference's own Triton kernel (vllm/attention/ops/prefix_prefill.py:630-755, from the
unmodified package in oracle/_ref/vllm_py) on identical inputs - the engine's 768-token budget
step (one to twelve chunks over prefixes of the trace's prompt lengths) and a long-prefix case.
flash_attn_varlen_func (what the engine uses for FIRST chunks, flash_attn.py:217-235) is timed on
the zero-prefix case for scale.  Development tool (run under gpurun).

    python tools/prefill_bench.py [--json out.json]
"""
import argparse
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = "cuda:0"


def load_triton():
    for base in (os.path.join(ROOT, "oracle", "_ref", "vllm_py"), "/root/reference"):
        path = os.path.join(base, "vllm", "attention", "ops", "prefix_prefill.py")
        if os.path.exists(path):
            spec = importlib.util.spec_from_file_location("ref_prefix_prefill", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.context_attention_fwd
    return None


def build(heads, hkv, D, bs, dtype, sub, ctx, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    B, T = len(sub), sum(sub)
    seq = [a + b for a, b in zip(sub, ctx)]
    u = lambda *s: ((torch.rand(*s, generator=g, device=DEV) * 2 - 1) * 0.5).to(dtype)
    q, k, v = u(T, heads, D), u(T, hkv, D), u(T, hkv, D)
    mb = max((s + bs - 1) // bs for s in seq)
    nb = B * mb + 64
    perm = np.random.default_rng(seed).permutation(nb)[:B * mb]
    bt = torch.tensor(perm, dtype=torch.int32).view(B, mb).to(DEV)
    x = 8
    kc, vc = u(nb, hkv, D // x, bs, x), u(nb, hkv, D, bs)
    start = torch.tensor([sum(sub[:i]) for i in range(B)], dtype=torch.int32, device=DEV)
    return dict(q=q, k=k, v=v, kc=kc, vc=vc, bt=bt, start=start,
                sl=torch.tensor(seq, dtype=torch.int32, device=DEV),
                cl=torch.tensor(ctx, dtype=torch.int32, device=DEV), max_sub=max(sub),
                cu=torch.tensor([0] + list(np.cumsum(sub)), dtype=torch.int32, device=DEV))


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from vllm_ltr_b200.prefix_prefill import context_attention_fwd as ours
    tri = load_triton()
    rng = np.random.default_rng(0)
    cases = []
    for name, heads, hkv, dtype in (("llama-2-7b", 32, 32, torch.float16),
                                    ("llama-3-8b", 32, 8, torch.bfloat16),
                                    ("70b-tp8", 8, 1, torch.bfloat16)):
        cases += [
            (name, heads, hkv, dtype, "1 chunk of 768 after a 1024-token prefix", [768], [1024]),
            (name, heads, hkv, dtype, "12 chunks of 64, prefixes 0..2047",
             [64] * 12, [0, 64, 128, 192, 256, 512, 1024, 1500, 2047, 300, 77, 900]),
            (name, heads, hkv, dtype, "3 chunks of 256 continuing prompts (prefix 256/512/768)",
             [256] * 3, [256, 512, 768]),
            (name, heads, hkv, dtype, "512-token budget left: 4 first chunks of 128 (no prefix)",
             [128] * 4, [0] * 4),
            (name, heads, hkv, dtype, "1 chunk of 64 after a 8000-token prefix", [64], [8000]),
        ]
    rows = []
    for name, heads, hkv, dtype, what, sub, ctx in cases:
        c = build(heads, hkv, 128, 16, dtype, sub, ctx)
        o1, o2 = torch.empty_like(c["q"]), torch.empty_like(c["q"])
        f_ours = lambda: ours(c["q"], c["k"], c["v"], o1, c["kc"], c["vc"], c["bt"], c["start"],
                              c["sl"], c["cl"], c["max_sub"], None, cache_has_new_tokens=True)
        # the cache must hold the chunk for the engine-path variant: write it once
        ours(c["q"], c["k"], c["v"], o1, c["kc"], c["vc"], c["bt"], c["start"], c["sl"], c["cl"],
             c["max_sub"], None, cache_has_new_tokens=False)
        t_ours = timeit(f_ours)
        t_tri = err = None
        if tri is not None:
            f_tri = lambda: tri(c["q"], c["k"], c["v"], o2, c["kc"], c["vc"], c["bt"], c["start"],
                                c["sl"], c["cl"], c["max_sub"])
            t_tri = timeit(f_tri)
            err = (o1.float() - o2.float()).abs().max().item()
        t_fa = None
        if all(x == 0 for x in ctx):
            from flash_attn import flash_attn_varlen_func
            f_fa = lambda: flash_attn_varlen_func(c["q"], c["k"], c["v"], c["cu"], c["cu"],
                                                  max(sub), max(sub), softmax_scale=128 ** -0.5,
                                                  causal=True)
            t_fa = timeit(f_fa)
        flops = sum(4 * heads * 128 * (s * cx + s * (s + 1) / 2) for s, cx in zip(sub, ctx))
        row = dict(model=name, case=what, ours_us=t_ours, triton_us=t_tri, flash_attn_us=t_fa,
                   speedup_vs_triton=(t_tri / t_ours if t_tri else None), max_abs_diff=err,
                   ours_tflops=flops / t_ours / 1e6)
        rows.append(row)
        print(f"{name:11s} {what:60s} ours {t_ours:8.1f} us  triton "
              f"{(t_tri if t_tri else float('nan')):8.1f} us  flash_attn "
              f"{(t_fa if t_fa else float('nan')):8.1f} us  x{(t_tri / t_ours if t_tri else 0):.2f}  "
              f"{row['ours_tflops']:.1f} TFLOP/s  diff {err if err is not None else float('nan'):.1e}",
              flush=True)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
