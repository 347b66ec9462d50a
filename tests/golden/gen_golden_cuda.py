#!/usr/bin/env python
"""Generates ref_cuda.npz from the REFERENCE's own CUDA kernels recompiled for sm_100a
(oracle/_ref/ref_cuda_C, built by `make -C oracle ref_cuda` from the sources under
/root/reference/csrc, unmodified).  Must run on a GPU box:

    gpurun -- 'python tests/golden/gen_golden_cuda.py gpurun_out/ref_cuda.npz'

and the result is committed as tests/golden/ref_cuda.npz.  Inputs are seeded and stored next to
the outputs (float32 upcasts, exact), so tests need neither the reference nor this module.
"""
import glob
import importlib.util
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

DEV = "cuda:0"
DT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def load():
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_cuda_C.*.so"))
    spec = importlib.util.spec_from_file_location("ref_cuda_C", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def f32(t):
    return t.detach().to("cpu", torch.float32).clone().numpy()


def main(path):
    ref = load()
    out = {}
    g = torch.Generator(device=DEV).manual_seed(2024)
    rnd = random.Random(2024)
    for name, dt in DT.items():
        x = (torch.randn(5, 2 * 128, generator=g, device=DEV) * 2).to(dt)
        out[f"act.{name}.x"] = f32(x)
        for act in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
            o = torch.empty(5, 128, dtype=dt, device=DEV)
            getattr(ref.ops, act)(o, x)
            out[f"{act}.{name}"] = f32(o)
        for act in ("gelu_new", "gelu_fast"):
            o = torch.empty_like(x)
            getattr(ref.ops, act)(o, x)
            out[f"{act}.{name}"] = f32(o)
        h = torch.randn(6, 512, generator=g, device=DEV).to(dt)
        r = torch.randn(6, 512, generator=g, device=DEV).to(dt)
        w = (torch.randn(512, generator=g, device=DEV) * 0.1 + 1).to(dt)
        o = torch.empty_like(h)
        ref.ops.rms_norm(o, h, w, 1e-5)
        out[f"rms.{name}.x"], out[f"rms.{name}.r"], out[f"rms.{name}.w"] = f32(h), f32(r), f32(w)
        out[f"rms.{name}.out"] = f32(o)
        h2, r2 = h.clone(), r.clone()
        ref.ops.fused_add_rms_norm(h2, r2, w, 1e-5)
        out[f"frms.{name}.out_x"], out[f"frms.{name}.out_r"] = f32(h2), f32(r2)
        # odd hidden size -> the reference's generic (non-vectorised) path
        h3 = torch.randn(3, 771, generator=g, device=DEV).to(dt)
        r3 = torch.randn(3, 771, generator=g, device=DEV).to(dt)
        w3 = (torch.randn(771, generator=g, device=DEV) * 0.1 + 1).to(dt)
        out[f"frms771.{name}.x"], out[f"frms771.{name}.r"], out[f"frms771.{name}.w"] = f32(h3), f32(r3), f32(w3)
        ref.ops.fused_add_rms_norm(h3, r3, w3, 1e-5)
        out[f"frms771.{name}.out_x"], out[f"frms771.{name}.out_r"] = f32(h3), f32(r3)
        cache = util.cos_sin_cache(64, 64, 512, dt).to(DEV)
        pos = torch.randint(0, 512, (9,), generator=g, device=DEV)
        q = torch.randn(9, 4 * 64, generator=g, device=DEV).to(dt)
        k = torch.randn(9, 2 * 64, generator=g, device=DEV).to(dt)
        out[f"rope.{name}.cache"], out[f"rope.{name}.pos"] = f32(cache), pos.cpu().numpy()
        out[f"rope.{name}.q"], out[f"rope.{name}.k"] = f32(q), f32(k)
        for neox in (1, 0):
            q2, k2 = q.clone(), k.clone()
            ref.ops.rotary_embedding(pos, q2, k2, 64, cache, bool(neox))
            out[f"rope.{name}.neox{neox}.q"], out[f"rope.{name}.neox{neox}.k"] = f32(q2), f32(k2)
        for kv in ("auto", "fp8"):
            for bs in (16, 32):
                hq, hkv, D, nb = 8, 2, 64, 48
                lens = [1, 15, 16, 17, 100, 513, 700]
                c = util.attention_case(len(lens), (hq, hkv), D, bs, dt, kv, True, 5, 0, DEV,
                                        context_lens=lens, num_blocks=nb)
                tag = f"attn.{name}.{kv}.bs{bs}"
                out[f"{tag}.q"] = f32(c["query"])
                if kv == "fp8":
                    out[f"{tag}.kc"] = c["key_cache"].cpu().numpy()
                    out[f"{tag}.vc"] = c["value_cache"].cpu().numpy()
                else:
                    out[f"{tag}.kc"], out[f"{tag}.vc"] = f32(c["key_cache"]), f32(c["value_cache"])
                out[f"{tag}.bt"] = c["block_tables"].cpu().numpy()
                out[f"{tag}.cl"] = c["context_lens"].cpu().numpy()
                out[f"{tag}.alibi"] = f32(c["alibi"])
                o1, _ = util.run_attention(ref.ops, c, "v1")
                o2, (ml, es, tmp) = util.run_attention(ref.ops, c, "v2")
                out[f"{tag}.v1"], out[f"{tag}.v2"] = f32(o1), f32(o2)
                # key/value rows written by reshape_and_cache (incl. the fp8 quantisation)
                key = (torch.randn(7, hkv, D, generator=g, device=DEV) * 3).to(dt)
                val = (torch.randn(7, hkv, D, generator=g, device=DEV) * 3).to(dt)
                slots = torch.tensor(rnd.sample(range(nb * bs), 7), dtype=torch.long, device=DEV)
                kc2, vc2 = c["key_cache"].clone(), c["value_cache"].clone()
                ref.cache_ops.reshape_and_cache(key, val, kc2, vc2, slots, kv, 1.0)
                out[f"{tag}.key"], out[f"{tag}.val"] = f32(key), f32(val)
                out[f"{tag}.slots"] = slots.cpu().numpy()
                if kv == "fp8":
                    out[f"{tag}.kc_out"], out[f"{tag}.vc_out"] = kc2.cpu().numpy(), vc2.cpu().numpy()
                else:
                    out[f"{tag}.kc_out"], out[f"{tag}.vc_out"] = f32(kc2), f32(vc2)
        # convert_fp8 over a dense sample of the type's range
        src = (torch.randn(4096, generator=g, device=DEV) *
               torch.logspace(-8, 5.5, 4096, device=DEV)).to(dt)
        dst = torch.empty(4096, dtype=torch.uint8, device=DEV)
        ref.cache_ops.convert_fp8(src, dst)
        back = torch.empty(256, dtype=dt, device=DEV)
        ref.cache_ops.convert_fp8(torch.arange(256, dtype=torch.uint8, device=DEV), back)
        out[f"fp8.{name}.src"], out[f"fp8.{name}.codes"] = f32(src), dst.cpu().numpy()
        out[f"fp8.{name}.back"] = f32(back)
    torch.cuda.synchronize()
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_cuda.npz"))
