#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE itself.  Run in the build container only
(needs /root/reference and oracle/_ref/ref_cpu_C built by `make -C oracle ref_cpu`):

    python tests/golden/gen_golden.py

Two sources, both the reference's own code, nothing of ours:
  ref_cpu_*   outputs of the reference's C++ CPU kernels (csrc/cpu/*.cpp compiled unmodified)
  ref_py_*    outputs of the reference's pure-torch layer formulas, imported from
              /root/reference: RMSNorm._forward (layers/layernorm.py:25-44),
              RotaryEmbedding._forward (layers/rotary_embedding.py:95-141),
              SiluAndMul/GeluAndMul/NewGELU/FastGELU._forward (layers/activation.py), and the
              attention oracle of tests/kernels/test_attention.py:57-113 (restated inline since
              that file imports xformers at module top).
Tensors are stored as float32 (exact upcasts of f16/bf16) plus the dtype name.
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
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)


def load_ref_cpu():
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_cpu_C.*.so"))
    spec = importlib.util.spec_from_file_location("ref_cpu_C", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def f32(t):
    return t.detach().to(torch.float32).clone().numpy()


DT = {"bfloat16": torch.bfloat16, "float32": torch.float32, "float16": torch.float16}


def kv_cache(nb, bs, hkv, D, dtype, g):
    x = 16 // torch.tensor([], dtype=dtype).element_size()
    sc = D ** -0.5
    k = ((torch.rand(nb, hkv, D // x, bs, x, generator=g) * 2 - 1) * sc).to(dtype)
    v = ((torch.rand(nb, hkv, D, bs, generator=g) * 2 - 1) * sc).to(dtype)
    return k, v


def gen_cpu(ref):
    out = {}
    g = torch.Generator().manual_seed(1234)
    rnd = random.Random(1234)
    for name in ("bfloat16", "float32"):
        dt = DT[name]
        # --- elementwise ---
        x = torch.randn(5, 2 * 96, generator=g).to(dt)
        for act in ("silu_and_mul", "gelu_and_mul", "gelu_tanh_and_mul"):
            o = torch.empty(5, 96, dtype=dt)
            getattr(ref.ops, act)(o, x)
            out[f"{act}.{name}.in"] = f32(x)
            out[f"{act}.{name}.out"] = f32(o)
        x1 = torch.randn(5, 96, generator=g).to(dt)
        for act in ("gelu_new", "gelu_fast"):
            o = torch.empty_like(x1)
            getattr(ref.ops, act)(o, x1)
            out[f"{act}.{name}.in"] = f32(x1)
            out[f"{act}.{name}.out"] = f32(o)
        # --- rms_norm / fused ---
        h = torch.randn(6, 256, generator=g).to(dt)
        r = torch.randn(6, 256, generator=g).to(dt)
        w = (torch.randn(256, generator=g) * 0.1 + 1).to(dt)
        o = torch.empty_like(h)
        ref.ops.rms_norm(o, h, w, 1e-5)
        out[f"rms.{name}.x"], out[f"rms.{name}.w"], out[f"rms.{name}.out"] = f32(h), f32(w), f32(o)
        h2, r2 = h.clone(), r.clone()
        ref.ops.fused_add_rms_norm(h2, r2, w, 1e-5)
        out[f"frms.{name}.r"] = f32(r)
        out[f"frms.{name}.out_x"], out[f"frms.{name}.out_r"] = f32(h2), f32(r2)
        # --- rotary (neox and gptj) ---
        from tests.util import cos_sin_cache
        cache = cos_sin_cache(64, 64, 512, dt)
        pos = torch.randint(0, 512, (9,), generator=g)
        q = torch.randn(9, 4 * 64, generator=g).to(dt)
        k = torch.randn(9, 2 * 64, generator=g).to(dt)
        out[f"rope.{name}.cache"], out[f"rope.{name}.pos"] = f32(cache), pos.numpy()
        out[f"rope.{name}.q"], out[f"rope.{name}.k"] = f32(q), f32(k)
        for neox in (True, False):
            q2, k2 = q.clone(), k.clone()
            ref.ops.rotary_embedding(pos, q2, k2, 64, cache, neox)
            out[f"rope.{name}.neox{int(neox)}.q"] = f32(q2)
            out[f"rope.{name}.neox{int(neox)}.k"] = f32(k2)
        # --- cache ops ---
        kc, vc = kv_cache(24, 16, 2, 64, dt, g)
        key = torch.randn(7, 2, 64, generator=g).to(dt)
        val = torch.randn(7, 2, 64, generator=g).to(dt)
        slots = torch.tensor(rnd.sample(range(24 * 16), 7), dtype=torch.long)
        out[f"cache.{name}.kc"], out[f"cache.{name}.vc"] = f32(kc), f32(vc)
        out[f"cache.{name}.key"], out[f"cache.{name}.val"] = f32(key), f32(val)
        out[f"cache.{name}.slots"] = slots.numpy()
        kc2, vc2 = kc.clone(), vc.clone()
        ref.cache_ops.reshape_and_cache(key, val, kc2, vc2, slots, "auto", 1.0)
        out[f"cache.{name}.kc_out"], out[f"cache.{name}.vc_out"] = f32(kc2), f32(vc2)
        mapping = {0: [5, 6], 3: [7], 20: [1]}
        ref.cache_ops.copy_blocks([kc2], [vc2], mapping)
        out[f"cache.{name}.kc_copied"], out[f"cache.{name}.vc_copied"] = f32(kc2), f32(vc2)
        # --- paged attention v1 / v2 (block 16 only on CPU) ---
        for hq, hkv in ((4, 4), (8, 2)):
            D, bs, nb = 64, 16, 64
            lens = [1, 15, 16, 17, 100, 513, 700]
            kc, vc = kv_cache(nb, bs, hkv, D, dt, g)
            qq = ((torch.rand(len(lens), hq, D, generator=g) * 2 - 1) * D ** -0.5).to(dt)
            mb = (max(lens) + bs - 1) // bs
            bt = torch.tensor([[rnd.randint(0, nb - 1) for _ in range(mb)] for _ in lens],
                              dtype=torch.int32)
            cl = torch.tensor(lens, dtype=torch.int32)
            alibi = torch.randn(hq, generator=g)
            tag = f"attn.{name}.h{hq}_{hkv}"
            out[f"{tag}.q"], out[f"{tag}.kc"], out[f"{tag}.vc"] = f32(qq), f32(kc), f32(vc)
            out[f"{tag}.bt"], out[f"{tag}.cl"], out[f"{tag}.alibi"] = bt.numpy(), cl.numpy(), f32(alibi)
            for al in (None, alibi):
                o = torch.empty_like(qq)
                ref.ops.paged_attention_v1(o, qq, kc, vc, hkv, D ** -0.5, bt, cl, bs, max(lens),
                                           al, "auto", 1.0)
                out[f"{tag}.v1.alibi{int(al is not None)}"] = f32(o)
                P = (max(lens) + 511) // 512
                o2 = torch.empty_like(qq)
                tmp = torch.empty(len(lens), hq, P, D, dtype=dt)
                es = torch.empty(len(lens), hq, P, dtype=torch.float32)
                ml = torch.empty_like(es)
                ref.ops.paged_attention_v2(o2, es, ml, tmp, qq, kc, vc, hkv, D ** -0.5, bt, cl,
                                           bs, max(lens), al, "auto", 1.0)
                out[f"{tag}.v2.alibi{int(al is not None)}"] = f32(o2)
    np.savez_compressed(os.path.join(HERE, "ref_cpu.npz"), **out)
    print("ref_cpu.npz:", len(out), "arrays")


def gen_py():
    """The reference's torch formulas (its own test oracles)."""
    from vllm.model_executor.layers.activation import FastGELU, GeluAndMul, NewGELU, SiluAndMul
    from vllm.model_executor.layers.layernorm import RMSNorm
    from vllm.model_executor.layers.rotary_embedding import get_rope
    out = {}
    g = torch.Generator().manual_seed(4321)
    for name, dt in DT.items():
        x = torch.randn(5, 2 * 96, generator=g).to(dt)
        out[f"act.{name}.x"] = f32(x)
        out[f"silu_and_mul.{name}"] = f32(SiluAndMul()._forward(x))
        out[f"gelu_and_mul.{name}"] = f32(GeluAndMul("none")._forward(x))
        out[f"gelu_tanh_and_mul.{name}"] = f32(GeluAndMul("tanh")._forward(x))
        out[f"gelu_new.{name}"] = f32(NewGELU()._forward(x))
        out[f"gelu_fast.{name}"] = f32(FastGELU()._forward(x))
        layer = RMSNorm(256).to(dtype=dt)
        layer.weight.data.copy_((torch.randn(256, generator=g) * 0.1 + 1).to(dt))
        h = (torch.randn(6, 256, generator=g) / 512).to(dt)
        r = (torch.randn(6, 256, generator=g) / 512).to(dt)
        out[f"rms.{name}.x"], out[f"rms.{name}.r"], out[f"rms.{name}.w"] = f32(h), f32(r), f32(layer.weight.data)
        out[f"rms.{name}.out"] = f32(layer._forward(h))
        a, b = layer._forward(h, r)
        out[f"frms.{name}.out_x"], out[f"frms.{name}.out_r"] = f32(a), f32(b)
        for neox in (True, False):
            rope = get_rope(64, 64, 512, 10000, neox).to(dtype=dt)
            pos = torch.randint(0, 512, (2, 5), generator=g)
            q = torch.randn(2, 5, 4 * 64, generator=g).to(dt)
            k = torch.randn(2, 5, 4 * 64, generator=g).to(dt)
            rq, rk = rope._forward(pos, q, k)
            tag = f"rope.{name}.neox{int(neox)}"
            out[f"{tag}.cache"] = f32(rope.cos_sin_cache.to(dt))
            out[f"{tag}.pos"], out[f"{tag}.q"], out[f"{tag}.k"] = pos.numpy(), f32(q), f32(k)
            out[f"{tag}.out_q"], out[f"{tag}.out_k"] = f32(rq), f32(rk)
    np.savez_compressed(os.path.join(HERE, "ref_py.npz"), **out)
    print("ref_py.npz:", len(out), "arrays")


if __name__ == "__main__":
    gen_cpu(load_ref_cpu())
    gen_py()
