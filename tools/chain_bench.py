#!/usr/bin/env python
"""In-step cost of the small decode-step kernels (development tool, run under gpurun).

A kernel timed alone pays a launch and a cold cache; inside the step it runs in a chain, with
programmatic dependent launch and (at N=1) from a CUDA graph.  This tool captures, per op, a graph
of `--n` consecutive launches on the buffers of one decoder layer and reports the time per launch,
next to the op's algorithmic bytes at the measured HBM peak - i.e. the floor each op adds to a layer.

    python tools/chain_bench.py [--model llama-2-7b] [--batch 256] [--n 64]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tools"))
from bench import build_inputs, trace_contexts  # noqa: E402

DEV = "cuda:0"


def graph_time(fn, n, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps / n * 1e3       # us per launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-2-7b,llama-3-8b,llama-3-8b-fp8kv")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from vllm_ltr_b200 import _C
    from vllm_ltr_b200._lib import tune
    from hotpath import SHAPES, DecodeHotPath, ModelShape, forward_decode
    peak = 6534.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    for name in args.model.split(","):
        s = ModelShape(**{**SHAPES[name].__dict__})
        s.dtype = torch.bfloat16
        ctx = trace_contexts(args.batch)
        nb = int(((ctx + 15) // 16).sum()) + 64
        hp = DecodeHotPath(s, nb, args.batch, DEV, layers=4)
        inp = {k: v.to(DEV) for k, v in build_inputs(ctx, 16, nb, 0).items()}
        T = args.batch
        q, k, v = hp.qkv[:T].split([s.q_heads * s.head_size, s.kv_heads * s.head_size,
                                    s.kv_heads * s.head_size], dim=-1)
        es = 2
        cs = 1 if s.kv_cache_dtype == "fp8" else 2
        layer = [0]

        def attn():
            i = layer[0] = (layer[0] + 1) % hp.layers
            forward_decode(q.view(T, s.q_heads, s.head_size), hp.key_caches[i], hp.value_caches[i],
                           inp["block_tables"], inp["context_lens"], int(ctx.max()),
                           s.kv_cache_dtype, s.kv_heads, hp.scale, None, 1.0,
                           out=hp.attn_out[:T].view(T, s.q_heads, s.head_size))

        def rms():
            _C.ops.fused_add_rms_norm(hp.o_proj_out[:T], hp.residual[:T], hp.norm_w[1], s.eps)

        def rope():
            _C.ops.rotary_embedding(inp["positions"], q, k, s.head_size, hp.cos_sin, True)

        clayer = [0]

        def reshape():
            # every decoder layer appends to its OWN cache: rotate, or the destination sectors
            # of the previous launch are still in L2 and the op looks twice as fast as it is
            i = clayer[0] = (clayer[0] + 1) % hp.layers
            _C.cache_ops.reshape_and_cache(k.view(T, s.kv_heads, s.head_size),
                                           v.view(T, s.kv_heads, s.head_size), hp.key_caches[i],
                                           hp.value_caches[i], inp["slot_mapping"],
                                           s.kv_cache_dtype, 1.0)

        def rope_cache():
            from vllm_ltr_b200.attention_backend import rotary_embedding_and_cache
            i = clayer[0] = (clayer[0] + 1) % hp.layers
            assert rotary_embedding_and_cache(inp["positions"], q, k, v, hp.key_caches[i],
                                              hp.value_caches[i], inp["slot_mapping"], s.head_size,
                                              hp.cos_sin, True, s.kv_cache_dtype)

        def silu():
            _C.ops.silu_and_mul(hp.act_out[:T], hp.gate_up[:T])

        ops = {
            "fused_add_rms_norm": (rms, T * s.hidden * es * 4 + s.hidden * es),
            "rotary_embedding": (rope, T * (s.q_heads + s.kv_heads) * s.head_size * es * 2
                                 + T * s.head_size * es + 8 * T),
            "reshape_and_cache": (reshape, T * 2 * s.kv_heads * s.head_size * (es + cs) + 8 * T),
            "rope + cache append (fused)": (rope_cache,
                                            T * 2 * s.q_heads * s.head_size * es
                                            + T * s.kv_heads * s.head_size * (3 * es + 2 * cs)
                                            + T * s.head_size * es + 16 * T),
            "silu_and_mul": (silu, T * s.inter * es * 3),
            "paged_attention": (attn, int(ctx.sum()) * 2 * s.kv_heads * s.head_size * cs
                                + 2 * T * s.q_heads * s.head_size * es),
        }

        def layer_small():
            rms(); rope(); reshape(); rms(); silu()

        def layer_full():
            rms(); rope(); reshape(); attn(); rms(); silu()

        print(f"== {name} B={T} (graph of {args.n} launches per op, PDL on / off)")
        for op, (fn, byts) in ops.items():
            res = []
            for pdl in (1, 0):
                tune("pdl", pdl)
                res.append(graph_time(fn, args.n))
            tune("pdl", 1)
            floor = byts / peak / 1e3
            print(f"   {op:28s} {res[0]:7.2f} us/launch (no PDL {res[1]:7.2f})   algorithmic "
                  f"{byts/1e6:8.2f} MB = {floor:6.2f} us at {peak:.0f} GB/s  ({floor/res[0]*100:4.0f}%)",
                  flush=True)
            rows.append(dict(model=name, op=op, us=res[0], us_nopdl=res[1], bytes=byts))
        def layer_small_fused():
            rms(); rope_cache(); rms(); silu()

        def layer_full_fused():
            rms(); rope_cache(); attn(); rms(); silu()

        for label, fn in (("layer without attention (5 kernels)", layer_small),
                          ("layer (6 kernels)", layer_full),
                          ("layer without attention, rope+cache fused (4)", layer_small_fused),
                          ("layer, rope+cache fused (5 kernels)", layer_full_fused)):
            res = []
            for pdl in (1, 0):
                tune("pdl", pdl)
                res.append(graph_time(fn, 16) )
            tune("pdl", 1)
            print(f"   {label:46s} {res[0]:7.2f} us (no PDL {res[1]:7.2f})", flush=True)
            rows.append(dict(model=name, op=label, us=res[0], us_nopdl=res[1]))
        del hp
        torch.cuda.empty_cache()
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
