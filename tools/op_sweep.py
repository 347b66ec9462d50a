#!/usr/bin/env python
"""Per-op micro-benchmark over the shapes of SURVEY.md section 8d: our kernels (through the
vllm._C mirror) next to the reference's CUDA kernels recompiled for sm_100a (oracle/_ref), with
achieved GB/s on ALGORITHMIC bytes.  Development tool (run under gpurun); bench.py is the contract.

    python tools/op_sweep.py [--ops attn,rms,silu,rope,cache] [--json out.json]
"""
import argparse
import glob
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import trace_contexts  # noqa: E402

DEV = "cuda:0"


def load_ref():
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_cuda_C.*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("ref_cuda_C", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_flush = None


def timeit(fn, iters=20, warmup=3, flush=True):
    """median device ms per call, L2 flushed between calls (writes 256 MiB)."""
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def attn_case(B, hq, hkv, dtype, kv, ctx_mode, seed=0):
    D, bs = 128, 16
    if ctx_mode == "trace":
        ctx = trace_contexts(B, seed)
    else:
        ctx = np.full(B, int(ctx_mode), dtype=np.int64)
    nblk = (ctx + bs - 1) // bs
    total = int(nblk.sum())
    nb = max(total + 16, 4096)
    es = 1 if kv == "fp8" else 2
    x = 16 // es
    cdt = torch.uint8 if kv == "fp8" else dtype
    g = torch.Generator(device=DEV).manual_seed(seed)
    if kv == "fp8":
        kc = torch.randint(0, 0x38, (nb, hkv, D // x, bs, x), generator=g, device=DEV, dtype=torch.uint8)
        vc = torch.randint(0, 0x38, (nb, hkv, D, bs), generator=g, device=DEV, dtype=torch.uint8)
    else:
        kc = ((torch.rand(nb, hkv, D // x, bs, x, generator=g, device=DEV) - 0.5) * 0.2).to(cdt)
        vc = ((torch.rand(nb, hkv, D, bs, generator=g, device=DEV) - 0.5) * 0.2).to(cdt)
    perm = np.random.default_rng(seed).permutation(nb)[:total]
    bt = np.zeros((B, int(nblk.max())), dtype=np.int32)
    o = 0
    for i in range(B):
        bt[i, :nblk[i]] = perm[o:o + nblk[i]]
        o += nblk[i]
    q = (torch.randn(B, hq, D, generator=g, device=DEV) * 0.1).to(dtype)
    byts = int(ctx.sum()) * 2 * hkv * D * es + 2 * B * hq * D * 2 + 4 * total + 4 * B
    return dict(q=q, kc=kc, vc=vc, bt=torch.from_numpy(bt).to(DEV),
                cl=torch.from_numpy(ctx.astype(np.int32)).to(DEV), mx=int(ctx.max()), hkv=hkv,
                kv=kv, bytes=byts, sum_ctx=int(ctx.sum()))


def run_attn(ops, c, version):
    q = c["q"]
    out = torch.empty_like(q)
    if version == "v1":
        return lambda: ops.paged_attention_v1(out, q, c["kc"], c["vc"], c["hkv"], 128 ** -0.5,
                                              c["bt"], c["cl"], 16, c["mx"], None, c["kv"], 1.0)
    B, H, D = q.shape
    P = (c["mx"] + 511) // 512
    tmp = torch.empty(B, H, P, D, dtype=q.dtype, device=DEV)
    es = torch.empty(B, H, P, dtype=torch.float32, device=DEV)
    ml = torch.empty_like(es)
    return lambda: ops.paged_attention_v2(out, es, ml, tmp, q, c["kc"], c["vc"], c["hkv"],
                                          128 ** -0.5, c["bt"], c["cl"], 16, c["mx"], None,
                                          c["kv"], 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="attn,rms,silu,rope,cache")
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="attn: cfg:B:mode:version, e.g. 8b-gqa-bf16:256:trace:v1")
    args = ap.parse_args()
    from vllm_ltr_b200 import _C
    ref = load_ref()
    peak = 6534.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []

    def report(op, desc, byts, ours_ms, ref_ms):
        gbs = byts / ours_ms / 1e6
        row = dict(op=op, shape=desc, bytes=byts, ours_ms=ours_ms, ours_gbs=gbs, frac=gbs / peak,
                   ref_ms=ref_ms, speedup=(ref_ms / ours_ms) if ref_ms else None)
        rows.append(row)
        print(f"{op:10s} {desc:46s} {byts/1e6:10.2f} MB  ours {ours_ms*1e3:9.1f} us {gbs:7.0f} GB/s "
              f"({gbs/peak*100:5.1f}%)  ref {ref_ms*1e3 if ref_ms else float('nan'):9.1f} us  "
              f"x{(ref_ms/ours_ms) if ref_ms else float('nan'):.2f}", flush=True)

    ops_sel = args.ops.split(",")
    if "attn" in ops_sel:
        cfgs = [("7b-mha-f16", 32, 32, torch.float16, "auto"),
                ("8b-gqa-bf16", 32, 8, torch.bfloat16, "auto"),
                ("8b-gqa-fp8", 32, 8, torch.bfloat16, "fp8"),
                ("70b-tp8-bf16", 8, 1, torch.bfloat16, "auto")]
        modes = [("trace", [1, 8, 64, 256]), ("512", [64, 256]), ("2048", [8, 64, 256]),
                 ("8192", [1, 8, 64])]
        if args.quick:
            modes = [("trace", [256]), ("2048", [64])]
        if args.only:
            oc, ob, om, ov = args.only.split(":")
            cfgs = [c for c in cfgs if c[0] == oc]
            modes = [(om, [int(ob)])]
        for name, hq, hkv, dt, kv in cfgs:
            for mode, Bs in modes:
                for B in Bs:
                    c = attn_case(B, hq, hkv, dt, kv, mode)
                    for ver in ("v1", "v2"):
                        if ver == "v2" and c["mx"] <= 512:
                            continue
                        if args.only and ver != ov:
                            continue
                        ours = timeit(run_attn(_C.ops, c, ver))
                        r = timeit(run_attn(ref.ops, c, ver)) if ref else None
                        report(f"attn_{ver}", f"{name} B={B} ctx={mode} sum={c['sum_ctx']}",
                               c["bytes"], ours, r)
                    del c
                    torch.cuda.empty_cache()
    g = torch.Generator(device=DEV).manual_seed(0)
    if "rms" in ops_sel:
        for T in (1, 64, 256, 768):
            for H in (4096, 8192):
                for dt in (torch.bfloat16,):
                    x = torch.randn(T, H, generator=g, device=DEV).to(dt)
                    r = torch.randn(T, H, generator=g, device=DEV).to(dt)
                    w = torch.randn(H, generator=g, device=DEV).to(dt)
                    byts = T * H * 2 * 4 + H * 2
                    ours = timeit(lambda: _C.ops.fused_add_rms_norm(x, r, w, 1e-5))
                    rr = timeit(lambda: ref.ops.fused_add_rms_norm(x, r, w, 1e-5)) if ref else None
                    report("fused_rms", f"T={T} H={H} bf16", byts, ours, rr)
    if "silu" in ops_sel:
        for T in (1, 64, 256, 768):
            for d in (11008, 14336, 3584):
                x = torch.randn(T, 2 * d, generator=g, device=DEV).to(torch.bfloat16)
                o = torch.empty(T, d, dtype=torch.bfloat16, device=DEV)
                ours = timeit(lambda: _C.ops.silu_and_mul(o, x))
                rr = timeit(lambda: ref.ops.silu_and_mul(o, x)) if ref else None
                report("silu_mul", f"T={T} d={d} bf16", T * d * 2 * 3, ours, rr)
    if "rope" in ops_sel:
        cache = torch.randn(8192, 128, generator=g, device=DEV).to(torch.bfloat16)
        for T in (1, 64, 256, 768):
            for hq, hkv in ((32, 32), (32, 8), (8, 1)):
                pos = torch.randint(0, 8192, (T,), generator=g, device=DEV)
                qkv = torch.randn(T, (hq + 2 * hkv) * 128, generator=g, device=DEV).to(torch.bfloat16)
                q, k, v = qkv.split([hq * 128, hkv * 128, hkv * 128], dim=-1)
                byts = T * (hq + hkv) * 128 * 2 * 2 + T * 128 * 2 + 8 * T
                ours = timeit(lambda: _C.ops.rotary_embedding(pos, q, k, 128, cache, True))
                rr = timeit(lambda: ref.ops.rotary_embedding(pos, q, k, 128, cache, True)) if ref else None
                report("rope", f"T={T} heads=({hq},{hkv}) bf16", byts, ours, rr)
    if "cache" in ops_sel:
        for T in (1, 64, 256, 768):
            for hkv, kv in ((32, "auto"), (8, "auto"), (8, "fp8"), (1, "auto")):
                nb, bs, D = 8192, 16, 128
                es = 1 if kv == "fp8" else 2
                x = 16 // es
                cdt = torch.uint8 if kv == "fp8" else torch.bfloat16
                kc = torch.zeros(nb, hkv, D // x, bs, x, dtype=cdt, device=DEV)
                vc = torch.zeros(nb, hkv, D, bs, dtype=cdt, device=DEV)
                qkv = torch.randn(T, (32 + 2 * hkv) * D, generator=g, device=DEV).to(torch.bfloat16)
                k = qkv[:, 32 * D:(32 + hkv) * D].view(T, hkv, D)
                v = qkv[:, (32 + hkv) * D:].view(T, hkv, D)
                slots = torch.randperm(nb * bs, generator=g, device=DEV)[:T]
                byts = T * 2 * hkv * D * (2 + es) + 8 * T
                ours = timeit(lambda: _C.cache_ops.reshape_and_cache(k, v, kc, vc, slots, kv, 1.0))
                rr = timeit(lambda: ref.cache_ops.reshape_and_cache(k, v, kc, vc, slots, kv, 1.0)) if ref else None
                report("reshape", f"T={T} kvh={hkv} {kv}", byts, ours, rr)
        # copy_blocks: 32 layers x 64 pairs, 7B block (16*32*128*2 = 128 KiB per K and V)
        L, nb = 32, 512
        ks = [torch.zeros(nb, 32, 16, 16, 8, dtype=torch.bfloat16, device=DEV) for _ in range(L)]
        vs = [torch.zeros(nb, 32, 128, 16, dtype=torch.bfloat16, device=DEV) for _ in range(L)]
        mapping = {i: [256 + i] for i in range(64)}
        byts = 64 * L * 2 * (16 * 32 * 128 * 2) * 2
        ours = timeit(lambda: _C.cache_ops.copy_blocks(ks, vs, mapping), iters=10)
        rr = timeit(lambda: ref.cache_ops.copy_blocks(ks, vs, mapping), iters=10) if ref else None
        report("copy_blocks", "32 layers x 64 pairs x 128 KiB", byts, ours, rr)
    if "swap" in ops_sel:
        # one preemption-sized swap-out: 32 layers x 64 blocks (7B block: 128 KiB per K and V) to
        # pinned host memory; ours = one launch for all layers, reference = its loop of
        # swap_blocks per layer and K/V (one cudaMemcpyAsync per block inside)
        L, nb, npairs = 32, 512, 64
        gpu = [torch.zeros(2, nb, 16 * 32 * 128, dtype=torch.bfloat16, device=DEV) for _ in range(L)]
        cpu = [torch.zeros(2, nb, 16 * 32 * 128, dtype=torch.bfloat16).pin_memory() for _ in range(L)]
        mapping = {i * 3: i for i in range(npairs)}
        byts = L * 2 * npairs * (16 * 32 * 128 * 2)

        def ref_loop():
            for i in range(L):
                ref.cache_ops.swap_blocks(gpu[i][0], cpu[i][0], mapping)
                ref.cache_ops.swap_blocks(gpu[i][1], cpu[i][1], mapping)

        def ours_loop():
            for i in range(L):
                _C.cache_ops.swap_blocks(gpu[i][0], cpu[i][0], mapping)
                _C.cache_ops.swap_blocks(gpu[i][1], cpu[i][1], mapping)
        import time as _t
        for name, fn in (("ours, one launch", lambda: _C.cache_ops.swap_blocks_layers(gpu, cpu, mapping)),
                         ("ours, per layer", ours_loop), ("reference loop", ref_loop if ref else None)):
            if fn is None:
                continue
            fn()
            torch.cuda.synchronize()
            t0 = _t.perf_counter()
            for _ in range(5):
                fn()
            t1 = _t.perf_counter()
            torch.cuda.synchronize()
            t2 = _t.perf_counter()
            print(f"swap_out   32 layers x 64 blocks x 128 KiB (K+V) = {byts/1e6:.0f} MB  {name:18s} "
                  f"host issue {(t1-t0)/5*1e3:7.2f} ms  end to end {(t2-t0)/5*1e3:7.2f} ms  "
                  f"{byts*5/(t2-t0)/1e9:6.1f} GB/s", flush=True)
            rows.append(dict(op="swap_out", shape=name, ms=(t2 - t0) / 5 * 1e3))
    if "sampler" in ops_sel:
        # next-token sampling: ours (one kernel) next to the reference's tensor program
        # (sampler.py:67-80, :333-357, :528-544) written with the same torch calls
        from vllm_ltr_b200.sampler import sample_rows
        for R, V, dt in ((256, 32000, torch.float16), (256, 128256, torch.bfloat16),
                         (64, 128256, torch.bfloat16), (8, 32000, torch.float16)):
            lg = (torch.randn(R, V, generator=g, device=DEV) * 3).to(dt)
            rnd = torch.ones(R, dtype=torch.uint8, device=DEV)

            def ref_prog():
                x = lg.clone()
                x.div_(torch.ones(R, 1, dtype=dt, device=DEV))
                probs = torch.softmax(x, dim=-1, dtype=torch.float)
                logprobs = torch.log_softmax(x, dim=-1, dtype=torch.float)
                q = torch.empty_like(probs).exponential_()
                tok = probs.div_(q).argmax(dim=1)
                lp = logprobs[torch.arange(R, device=DEV), tok]
                return tok, lp, (logprobs > lp[:, None]).long().sum(1).add_(1)
            ours = timeit(lambda: sample_rows(lg, None, rnd, None, seed=1, offset=0))
            rr = timeit(ref_prog)
            report("sampler", f"rows={R} vocab={V} {str(dt).split('.')[-1]} random", R * V * lg.element_size(),
                   ours, rr)
    if "gemm" in ops_sel:
        # the predictor's dense layers: ours (tcgen05) next to cuBLAS (torch.nn.functional.linear)
        from vllm_ltr_b200 import predictor as P
        for M in (245, 2048):
            for N, K, relu, resid in ((2304, 768, False, False), (768, 768, False, True),
                                      (3072, 768, True, False), (768, 3072, False, True),
                                      (4096, 1024, True, False), (1024, 4096, False, True)):
                a = torch.randn(M, K, generator=g, device=DEV).half()
                w = (torch.randn(N, K, generator=g, device=DEV) * 0.05).half()
                b = torch.randn(N, generator=g, device=DEV).half()
                r = torch.randn(M, N, generator=g, device=DEV).half() if resid else None
                ours = timeit(lambda: P.gemm(a, w, bias=b, resid=r, relu=relu), flush=False)

                def cublas():
                    y = torch.nn.functional.linear(a, w, b)
                    if relu:
                        y = torch.relu(y)
                    if r is not None:
                        y = y + r
                    return y
                rr = timeit(cublas, flush=False)
                fl = 2.0 * M * N * K
                print(f"gemm       M={M} N={N} K={K} relu={int(relu)} resid={int(resid)}  ours {ours*1e3:8.1f} us "
                      f"{fl/ours/1e9:7.1f} TFLOP/s   cuBLAS+epilogue {rr*1e3:8.1f} us {fl/rr/1e9:7.1f} TFLOP/s  x{rr/ours:.2f}",
                      flush=True)
                rows.append(dict(op="gemm", shape=f"M={M} N={N} K={K}", ours_ms=ours, ref_ms=rr,
                                 tflops=fl / ours / 1e9))
    if "predictor" in ops_sel:
        from vllm_ltr_b200 import predictor as P
        for kind in ("125m", "350m"):
            cfg, sd = P.dummy_checkpoint(kind)
            pred = P.OPTPredictor(cfg, sd, device=DEV)
            rng = np.random.default_rng(0)
            for desc, lens in (("1x245", [245]), ("8x256", [256] * 8), ("1x2048", [2048]),
                               ("14x~146", [146] * 14)):
                prompts = [rng.integers(1000, 49000, L).tolist() for L in lens]
                ms = timeit(lambda: pred.logits(prompts), iters=10, flush=False)
                toks = sum(lens)
                print(f"predictor  opt-{kind} prompts={desc} tokens={toks}  {ms*1e3:9.1f} us  "
                      f"{toks/ms/1e3:8.1f} Mtok/s", flush=True)
                rows.append(dict(op="predictor", shape=f"{kind} {desc}", ours_ms=ms))
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
