#!/usr/bin/env python
"""bench.py - decode hot-path benchmark (driver contract: one JSON line on stdout from rank 0).

Metric (BASELINE.json): output tokens/s on a ShareGPT-shaped trace.  One "step" = one engine
decode step of the hot path for a batch of B running sequences (one new token each): for every
decoder layer  fused_add_rms_norm -> rotary_embedding -> reshape_and_cache -> paged_attention ->
fused_add_rms_norm -> silu_and_mul, then the final norm (tools/hotpath.py; in our arm the rotary
embedding and the cache append are one launch, csrc/rope_cache.cu; the cuBLAS
GEMMs between them are library code outside this path and are NOT executed or counted).
Contexts are the running contexts of a synthetic ShareGPT-shaped trace.

  value     tokens/s with every input already resident in HBM (CUDA events, max over ranks);
            the step is replayed from a CUDA graph (DecodeHotPath.capture) at every N
  e2e       the same step driven from HOST buffers: per step the engine's host-built inputs
            (positions, slot_mapping, block_tables, context_lens, token hidden rows) are copied
            from pinned memory, and the final-norm output is read back, inside the timed region
  roofline  paged attention (the dominant kernel): algorithmic bytes per call / mean call time,
            every call bracketed by CUDA events in an eager pass of the same K steps
  cpu_baseline / --impl reference: the reference's own CPU kernels (csrc/cpu, built into
            oracle/_ref by oracle/Makefile) driven through the same call sequence on host cores.

  ref_cuda  (N=1) the same step, pools and inputs driven through the REFERENCE's CUDA kernels
            recompiled for sm_100a (oracle/_ref/ref_cuda_C: the "reference CUDA build" target of
            BASELINE.md section 3.1), replayed from a CUDA graph like ours, with its own clocks;
            `models` repeats the ours / ref_cuda pair for Llama-3-8B and Llama-3-8B + e5m2 KV
  with_gemms  the step with the four cuBLAS projections of every layer between the ops
  trace     (N=1) the reference's real metric: the UNMODIFIED reference engine
            (oracle/_ref/vllm_py) runs the synthetic ShareGPT trace of BASELINE.md section 4 -
            10 000 requests at t=0, "opt-xxx-synthetic((-1,-1,1000),)", chunked prefill, eager,
            dummy-weight Llama-2-7B + OPT-125M predictor - once with the reference's CUDA kernels
            and stock scheduler / predictor engine and once with vllm_ltr_b200.install() +
            install_engine(); output tokens/s and p50 request latency of both
            (tools/engine_trace.py, one subprocess per arm)
  trace_poisson  (N=1) the same two engines with the trace's requests ARRIVING over time
            (Gamma(cv=1) gaps at --trace-poisson-rate requests/s for --trace-poisson-time s,
            benchmark_serving_real.py:159-176 / bench-lmsys.sh), schedule "opt-xxx", every
            request run to its end; output tokens/s and p50 latency from arrival

N > 1 (torchrun): tensor-parallel shards of the same model (q/kv heads and MLP width divided,
two exchanges of [B, hidden] per layer, fused with the residual add + RMSNorm that follows),
scaling = "strong"; the step replays from a CUDA graph on every rank in lockstep.  Before
timing, the fused exchange is checked bit-exactly against NCCL on integer-valued inputs
(`tp_parity`); at 8 GPUs the Llama-2-70B TP8 shapes are measured as well (`config_70b`).
"""
import argparse
import glob
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TOOLS = os.path.join(ROOT, "tools")      # bench / test drivers (hotpath.py, engine_trace.py)
if TOOLS not in sys.path:
    sys.path.insert(1, TOOLS)


# ---------------------------------------------------------------------------------------------
# workload: running contexts of the synthetic ShareGPT-shaped trace
# ---------------------------------------------------------------------------------------------
def trace_contexts(batch: int, seed: int = 0, max_ctx: int = 8192):
    rng = np.random.default_rng(seed)
    prompt = np.clip(np.rint(np.exp(rng.normal(5.0, 1.0, batch))), 4, 2048).astype(np.int64)
    out = np.clip(np.rint(np.exp(rng.normal(5.3, 1.1, batch))), 4, max_ctx - prompt).astype(np.int64)
    done = (rng.random(batch) * out).astype(np.int64)       # tokens generated so far
    ctx = np.minimum(prompt + done + 1, max_ctx)            # context seen by this decode step
    return ctx


def build_inputs(ctx: np.ndarray, block_size: int, num_blocks: int, seed: int):
    """Host-side step inputs exactly as ModelRunner._prepare_decode builds them
    (vllm/worker/model_runner.py:424-548): positions, slot_mapping, padded block_tables,
    context_lens.  Block ids are a random permutation of the pool (worst-case locality)."""
    rng = np.random.default_rng(seed + 1)
    B = len(ctx)
    nblk = (ctx + block_size - 1) // block_size
    assert int(nblk.sum()) <= num_blocks
    perm = rng.permutation(num_blocks)[: int(nblk.sum())]
    mb = int(nblk.max())
    bt = np.zeros((B, mb), dtype=np.int32)
    o = 0
    for i in range(B):
        bt[i, : nblk[i]] = perm[o:o + nblk[i]]
        o += nblk[i]
    pos = (ctx - 1).astype(np.int64)
    slot = bt[np.arange(B), pos // block_size].astype(np.int64) * block_size + pos % block_size
    return dict(positions=torch.from_numpy(pos), slot_mapping=torch.from_numpy(slot),
                block_tables=torch.from_numpy(bt),
                context_lens=torch.from_numpy(ctx.astype(np.int32)))


def attention_bytes(ctx, shape, block_size):
    """Algorithmic bytes of one paged-attention call."""
    s_cache = 1 if shape.kv_cache_dtype == "fp8" else torch.tensor([], dtype=shape.dtype).element_size()
    s_act = torch.tensor([], dtype=shape.dtype).element_size()
    B = len(ctx)
    kv = int(ctx.sum()) * 2 * shape.kv_heads * shape.head_size * s_cache
    qo = 2 * B * shape.q_heads * shape.head_size * s_act
    tbl = 4 * int(((ctx + block_size - 1) // block_size).sum()) + 4 * B
    return kv + qo + tbl


def step_bytes(ctx, shape, block_size, layers):
    """Algorithmic bytes of the whole step (all in-scope ops)."""
    s = torch.tensor([], dtype=shape.dtype).element_size()
    s_cache = 1 if shape.kv_cache_dtype == "fp8" else s
    T, H, D = len(ctx), shape.hidden, shape.head_size
    per_layer = (attention_bytes(ctx, shape, block_size)
                 + 2 * (T * H * s * 4 + H * s)                                  # 2 fused norms
                 + T * (shape.q_heads + shape.kv_heads) * D * s * 2 + T * D * s + 8 * T   # rope
                 + T * 2 * shape.kv_heads * D * (s + s_cache) + 8 * T           # reshape_and_cache
                 + T * shape.inter * s * 3)                                     # silu_and_mul
    return per_layer * layers + T * H * s * 4 + H * s


# ---------------------------------------------------------------------------------------------
# the LTR scheduling tick that runs once per engine step
# ---------------------------------------------------------------------------------------------
def tick_workload(n_live: int, prompt_len: int, seed: int = 0):
    """fp16-quantised scores of the live requests (what logits[:,0].tolist() yields) and the
    token ids of the prompt that arrives this step (steady state of the trace: completions per
    step ~ B / mean output length < 1, so one new prompt of the trace's mean length per step)."""
    rng = np.random.default_rng(seed + 7)
    scores = rng.normal(0, 1, n_live).astype(np.float16).astype(np.float64)
    prompt = rng.integers(1000, 49000, prompt_len).tolist()
    return scores, prompt


class GpuTick:
    """a8 + a10 on a side stream: score the newly arrived prompt with the OPT-125M predictor and
    order all live requests by score; the scheduler thread waits only on that stream."""

    def __init__(self, dev, n_live, prompt_len):
        from vllm_ltr_b200.ltr_scheduler import GpuRanker
        from vllm_ltr_b200.predictor import OPTPredictor, dummy_checkpoint
        self.scores, self.prompt = tick_workload(n_live, prompt_len)
        cfg, sd = dummy_checkpoint("125m", seed=0)
        self.pred = OPTPredictor(cfg, sd, device=dev)
        self.ranker = GpuRanker()
        self.stream = torch.cuda.Stream(device=dev, priority=-1)   # ahead of the model stream
        self.last = None

    def __call__(self):
        with torch.cuda.stream(self.stream):
            s = self.pred.scores([self.prompt])          # D2H of the score: syncs the side stream
            self.scores[0] = s[0]
            self.last = self.ranker.order(self.scores)   # H2D keys, rank kernel, D2H permutation
        return self.last


class CpuTick:
    """The same tick the way the reference does it on the host: Python's sorted over the live
    requests (scheduler.py:998) and an fp32 forward of the predictor architecture (HF
    OPTForSequenceClassification, the class train/trainer.py saves) on the host cores."""

    def __init__(self, n_live, prompt_len):
        import transformers
        self.scores, self.prompt = tick_workload(n_live, prompt_len)
        cfg = transformers.OPTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                                     ffn_dim=3072, word_embed_proj_dim=768, vocab_size=50272,
                                     max_position_embeddings=2048, num_labels=1,
                                     do_layer_norm_before=True)
        torch.manual_seed(0)
        self.model = transformers.OPTForSequenceClassification(cfg).eval()
        self.ids = torch.tensor([self.prompt])

    def __call__(self):
        with torch.no_grad():
            s = float(self.model(input_ids=self.ids).logits[0, 0].to(torch.float16))
        sc = self.scores.tolist()
        sc[0] = s
        return sorted(range(len(sc)), key=lambda i: -sc[i])


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference CPU arm
# ---------------------------------------------------------------------------------------------
def load_ref_cpu():
    """The reference's own CPU kernels (csrc/cpu/*.cpp), prebuilt into oracle/_ref by
    oracle/Makefile.  Returns (module, kind)."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        flags = ""
    stems = (["ref_cpu_C"] if "avx512_bf16" in flags else []) + ["ref_cpu_nobf16_C"]
    if "avx512f" not in flags:
        stems = []
    for stem in stems:
        hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", stem + ".*.so"))
        if hits:
            spec = importlib.util.spec_from_file_location(stem, hits[0])
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod, "reference"
    return None, "port"


class OraclePort:
    """oracle/kernels.py dressed as a vllm._C module (used only when oracle/_ref is missing)."""

    def __init__(self):
        from oracle import kernels as K
        port = self

        class Ops:
            @staticmethod
            def rms_norm(out, x, w, eps):
                out.copy_(K.rms_norm(x, w, eps))

            @staticmethod
            def fused_add_rms_norm(x, r, w, eps):
                a, b = K.fused_add_rms_norm(x, r, w, eps)
                x.copy_(a)
                r.copy_(b)

            @staticmethod
            def rotary_embedding(pos, q, k, hs, cache, neox):
                a, b = K.rotary_embedding(pos, q, k, hs, cache, neox)
                q.copy_(a)
                k.copy_(b)

            @staticmethod
            def silu_and_mul(out, x):
                out.copy_(K.activation(x, "silu_and_mul"))

            @staticmethod
            def paged_attention_v1(out, q, kc, vc, hkv, scale, bt, cl, bs, mx, alibi, kvd, kvs):
                out.copy_(K.paged_attention(q, kc, vc, bt, cl, scale, alibi, kvd).to(out.dtype))

            @staticmethod
            def paged_attention_v2(out, es, ml, tmp, q, kc, vc, hkv, scale, bt, cl, bs, mx, alibi,
                                   kvd, kvs):
                # same result; the partition scratch is an implementation detail of the kernels
                out.copy_(K.paged_attention(q, kc, vc, bt, cl, scale, alibi, kvd).to(out.dtype))

        class CacheOps:
            @staticmethod
            def reshape_and_cache(k, v, kc, vc, slots, kvd, kvs):
                K.reshape_and_cache(k, v, kc, vc, slots, kvd)

        self.ops, self.cache_ops = Ops, CacheOps


def cpu_reference_run(args, shape, ctx, steps, warmup, sample_layers, sample_batch):
    """Times the reference CPU kernels on a bounded sample: `sample_batch` sequences of the same
    trace x `sample_layers` layers per step, scaled to the full layer count.  Returns
    (tokens/s, ms per full step, description, cores, kind)."""
    from hotpath import DecodeHotPath, ModelShape
    mod, kind = load_ref_cpu()
    if mod is None:
        mod = OraclePort()
        sample_batch = min(sample_batch, 8)
        sample_layers = 1
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # torchrun exports OMP_NUM_THREADS=1; the reference kernels read the OpenMP runtime directly
    import ctypes
    for so in ("libgomp.so.1",):
        try:
            ctypes.CDLL(so).omp_set_num_threads(cores)
        except OSError:
            pass
    cshape = ModelShape(**{**shape.__dict__})
    if cshape.dtype == torch.float16:
        cshape.dtype = torch.bfloat16        # csrc/cpu implements f32 and bf16 only
    cshape.kv_cache_dtype = "auto"            # and has no fp8 cache
    sctx = ctx[:sample_batch]
    bs = cshape.block_size
    need = int(((sctx + bs - 1) // bs).sum())
    hp = DecodeHotPath(cshape, need + 8, len(sctx), "cpu", seed=0, layers=sample_layers, C=mod)
    inp = build_inputs(sctx, bs, need + 8, 0)
    hidden = torch.randn(len(sctx), cshape.hidden).to(cshape.dtype)
    mx = int(sctx.max())

    tick = CpuTick(args.live_requests, args.tick_prompt_len) if args.tick else None

    def one():
        hp.step(hidden, inp["positions"], inp["slot_mapping"], inp["block_tables"],
                inp["context_lens"], mx)

    for _ in range(warmup):
        one()
    for _ in range(2 if tick else 0):      # first calls pay thread-pool / allocator start-up
        tick()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    tick_s = 0.0
    if tick:
        t0 = time.perf_counter()
        for _ in range(steps):
            tick()
        tick_s = (time.perf_counter() - t0) / steps
    full = dt * shape.layers / sample_layers + tick_s
    tps = len(sctx) / full
    desc = (f"{len(sctx)} of {len(ctx)} sequences of the trace x {sample_layers} of {shape.layers} "
            f"layers per step, time scaled by {shape.layers}/{sample_layers}; "
            f"csrc/cpu kernels ({'oracle/_ref' if kind == 'reference' else 'oracle port'}), "
            f"{str(cshape.dtype).split('.')[-1]}, OMP threads={cores}"
            + (f"; + LTR tick per step ({tick_s * 1e3:.1f} ms: Python sorted over "
               f"{args.live_requests} live requests + HF OPT-125M fp32 forward of one "
               f"{args.tick_prompt_len}-token prompt)" if tick else ""))
    return tps, full * 1e3, desc, cores, kind


# ---------------------------------------------------------------------------------------------
# secondary arms (N=1): reference CUDA kernels on the same step, other models, the engine trace
# ---------------------------------------------------------------------------------------------
def load_ref_cuda():
    """The reference's CUDA kernels recompiled unmodified for sm_100a (oracle/_ref/ref_cuda_C,
    built by oracle/Makefile) - a checker / baseline, never on the product path."""
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_cuda_C.*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("ref_cuda_C", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _sync():
    torch.cuda.synchronize()


def time_step_graph(hp, devin, max_ctx, steps):
    """ms per step of the graph-replayed decode step (inputs resident, no tick)."""
    g, _ = hp.capture(devin["hidden"], devin["positions"], devin["slot_mapping"],
                      devin["block_tables"], devin["context_lens"], max_ctx)
    for _ in range(3):
        g.replay()
    _sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    _sync()
    return e0.elapsed_time(e1) / steps


def time_attention_b2b(hp, devin, max_ctx, batch, steps):
    for _ in range(2):
        hp.attention_only(devin["block_tables"], devin["context_lens"], max_ctx, batch)
    _sync()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(steps):
        hp.attention_only(devin["block_tables"], devin["context_lens"], max_ctx, batch)
    a1.record()
    _sync()
    return a0.elapsed_time(a1) / (steps * hp.layers)


def step_pair(name, batch, steps, dev, ref, peak, clocks, ours=True, layers=None):
    """ours / ref_cuda on the decode step of model `name`: same pools, same inputs, graph replay."""
    from hotpath import SHAPES, DecodeHotPath, ModelShape
    shape = ModelShape(**{**SHAPES[name].__dict__})
    if shape.dtype == torch.float16:
        shape.dtype = torch.bfloat16
    if layers:
        shape.layers = layers
    ctx = trace_contexts(batch)
    bs = shape.block_size
    nb = int(((ctx + bs - 1) // bs).sum()) + 64
    devin = {k: v.to(dev) for k, v in build_inputs(ctx, bs, nb, 0).items()}
    g = torch.Generator().manual_seed(0)
    devin["hidden"] = (torch.randn(batch, shape.hidden, generator=g) * 0.5).to(shape.dtype).to(dev)
    max_ctx = int(ctx.max())
    abytes = attention_bytes(ctx, shape, bs)
    res = {"model": name, "batch": batch, "layers": shape.layers,
           "attention_algorithmic_bytes": abytes}
    arms = ([("ours", None)] if ours else []) + ([("ref_cuda", ref)] if ref is not None else [])
    for label, mod in arms:
        hp = DecodeHotPath(shape, nb, batch, dev, C=mod)
        for _ in range(3):
            hp.step(devin["hidden"], devin["positions"], devin["slot_mapping"],
                    devin["block_tables"], devin["context_lens"], max_ctx)
        _sync()
        t0 = time.time()
        ms = time_step_graph(hp, devin, max_ctx, steps)
        a_ms = time_attention_b2b(hp, devin, max_ctx, batch, steps)
        t1 = time.time()
        res[label] = {"ms_per_step": ms, "tokens_per_s": batch / (ms * 1e-3),
                      "attention_launch_ms": a_ms,
                      "attention_gbs": abytes / (a_ms * 1e-3) / 1e9,
                      "attention_roofline_frac": abytes / (a_ms * 1e-3) / 1e9 / peak,
                      "clocks": clocks.summary(t0, t1)}
        del hp
        torch.cuda.empty_cache()
    if "ours" in res and "ref_cuda" in res:
        res["speedup_vs_ref_cuda"] = res["ref_cuda"]["ms_per_step"] / res["ours"]["ms_per_step"]
    return res


def run_trace_arm(arm, args, clocks, timeout_s, poisson=False):
    """One arm of the engine-level trace in its own process (vllm._C is bound per process).
    `poisson`: requests arrive at --trace-poisson-rate for --trace-poisson-time seconds and every
    one runs to its end (no synthetic stop rule); otherwise all requests are queued at t=0."""
    out = os.path.join("/tmp", f"vltr_trace_{arm}_{os.getpid()}.json")
    cmd = [sys.executable, os.path.join(TOOLS, "engine_trace.py"), "--arm", arm,
           "--model", args.trace_model, "--json", out]
    if poisson:
        n = int(args.trace_poisson_rate * args.trace_poisson_time * 1.5) + 16
        cmd += ["--num-prompts", str(n), "--stop-after", "0",
                "--request-rate", str(args.trace_poisson_rate),
                "--request-time", str(args.trace_poisson_time),
                "--max-output", str(args.trace_poisson_max_output)]
    else:
        cmd += ["--num-prompts", str(args.trace_prompts), "--stop-after", str(args.trace_stop)]
    t0 = time.time()
    try:
        r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=timeout_s)
    except subprocess.TimeoutExpired:
        return {"error": f"timeout after {timeout_s:.0f} s"}
    t1 = time.time()
    if r.returncode != 0 or not os.path.exists(out):
        return {"error": f"rc={r.returncode}", "tail": r.stdout[-1500:]}
    with open(out) as f:
        d = json.load(f)
    os.unlink(out)
    s = d["summary"]
    keep = ("tokens_per_s", "p50_latency_s", "p90_latency_s", "mean_latency_s", "requests_per_s",
            "num_requests", "finished", "generated_tokens", "elapsed_s", "steps", "schedule_s",
            "execute_model_s", "predictor_s", "other_host_s", "add_requests_s", "arrivals")
    res = {k: s[k] for k in keep if k in s}
    res["engine_init_s"] = d["engine"]["init_s"]
    res["num_gpu_blocks"] = d["engine"]["num_gpu_blocks"]
    res["wall_s"] = t1 - t0
    res["clocks"] = clocks.summary(t1 - s["elapsed_s"], t1)
    if "installed" in d["engine"]:
        res["installed"] = sorted(k for k, v in d["engine"]["installed"].items()
                                  if not v.startswith("skipped"))
    return res


def measure_70b(args, dev, rank, world, peak):
    """BASELINE.json config 5 at step level: Llama-2-70B shapes sharded TP=`world` (8 q / 1 kv
    head and inter/8 per GPU, 80 layers, 160 fused exchanges of [B, 8192] per step), same trace
    batch, graph replay on every rank in lockstep, max over ranks."""
    import torch.distributed as dist
    from hotpath import SHAPES, DecodeHotPath, ModelShape
    from vllm_ltr_b200.parallel import FusedTPComm
    full = ModelShape(**{**SHAPES["llama-2-70b"].__dict__})
    full.dtype = torch.bfloat16
    shape = full.shard(world)
    ctx = trace_contexts(args.batch)
    bs = shape.block_size
    nb = int(((ctx + bs - 1) // bs).sum()) + 64
    devin = {k: v.to(dev) for k, v in build_inputs(ctx, bs, nb, 0).items()}
    g = torch.Generator().manual_seed(0)
    devin["hidden"] = (torch.randn(args.batch, shape.hidden, generator=g) * 0.5).to(shape.dtype).to(dev)
    max_ctx = int(ctx.max())
    hp = DecodeHotPath(shape, nb, args.batch, dev, seed=rank)
    hp.attach_fused_tp(FusedTPComm(args.batch, shape.hidden, shape.dtype, dev))
    parity = tp_parity_check(hp.comm, dev, rank, world, args.batch, shape.hidden, shape.dtype)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        hp.step(devin["hidden"], devin["positions"], devin["slot_mapping"], devin["block_tables"],
                devin["context_lens"], max_ctx)
    barrier()
    graph, _ = hp.capture(devin["hidden"], devin["positions"], devin["slot_mapping"],
                          devin["block_tables"], devin["context_lens"], max_ctx)
    for _ in range(3):
        graph.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        graph.replay()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    a_ms = time_attention_b2b(hp, devin, max_ctx, args.batch, args.steps)
    barrier()
    t = torch.tensor([ms, a_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, a_ms = float(t[0]), float(t[1])
    abytes = attention_bytes(ctx, shape, bs)
    del hp, graph
    torch.cuda.empty_cache()
    return {"workload": f"llama-2-70b TP{world} decode hot path, B={args.batch}, {shape.layers} "
                        "layers, fused exchange+norm, CUDA-graph replay, no tick",
            "ms_per_step": ms, "tokens_per_s": args.batch / (ms * 1e-3), "tp_parity": parity,
            "attention_launch_ms": a_ms, "attention_algorithmic_bytes": abytes,
            "attention_roofline_frac": abytes / (a_ms * 1e-3) / 1e9 / peak,
            "attention_share_of_step": a_ms * shape.layers / ms}


def tp_parity_check(comm, dev, rank, world, T, H, dtype):
    """Integer-valued exactness of the fused exchange against NCCL (the reference's own test of
    its custom all-reduce: tests/distributed/test_custom_all_reduce.py:30-51): the plain sum must
    equal dist.all_reduce bit for bit, the fused residual too, and the normalised rows must
    equal our own norm kernel applied to the NCCL sum.  Returns "exact" or raises."""
    import torch.distributed as dist
    from vllm_ltr_b200 import _C
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randint(-4, 5, (T, H), generator=g, device=dev).to(dtype)
    g0 = torch.Generator(device=dev).manual_seed(7)
    resid = torch.randint(-4, 5, (T, H), generator=g0, device=dev).to(dtype)
    w = (torch.randn(H, generator=g0, device=dev) * 0.1 + 1).to(dtype)
    keep = [t[:T].clone() for t in (comm.x[0], comm.x[1], comm.resid)]
    s = x.clone()
    dist.all_reduce(s)
    want_res = resid.clone()
    want_out = s.clone()
    _C.ops.fused_add_rms_norm(want_out, want_res, w, 1e-5)     # residual += sum; out = norm
    comm.x[1][:T].copy_(x)
    torch.cuda.synchronize()
    dist.barrier()
    plain = comm.allreduce(1, T).clone()
    comm.x[0][:T].copy_(x)
    comm.resid[:T].copy_(resid)
    torch.cuda.synchronize()
    dist.barrier()
    out = comm.allreduce_add_rms_norm(0, w, 1e-5, T).clone()
    torch.cuda.synchronize()
    # the residual is replicated or owner-local (row % world == rank), as the communicator says
    mine = (torch.ones(T, dtype=torch.bool, device=dev) if getattr(comm, "replicate_residual", True)
            else torch.arange(T, device=dev) % world == rank)
    ok = torch.tensor([int(torch.equal(plain, s)),
                       int(torch.equal(comm.resid[:T][mine], want_res[mine])),
                       int(torch.equal(out, want_out))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    for t, k in zip((comm.x[0], comm.x[1], comm.resid), keep):
        t[:T].copy_(k)
    torch.cuda.synchronize()
    dist.barrier()
    if not bool(ok.all()):
        raise RuntimeError(f"fused TP exchange differs from NCCL: sum/resid/norm exact = {ok.tolist()}")
    return "exact"


# ---------------------------------------------------------------------------------------------
def main():
    # the driver reads ONE JSON line from stdout: libraries that print there (NCCL's version
    # banner) are pointed at stderr, the line goes to the real stdout at the end
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-2-7b")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=None, help="debug: fewer layers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-layers", type=int, default=2)
    ap.add_argument("--cpu-sample-batch", type=int, default=256)
    ap.add_argument("--tp-exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused peer-memory all-reduce+norm kernel, or NCCL all_reduce + norm")
    ap.add_argument("--launch", default="graph", choices=["graph", "eager"],
                    help="replay the step from a CUDA graph (default) or issue it kernel by kernel")
    ap.add_argument("--with-gemms", action="store_true",
                    help="N=1, context only: additionally time the step with the four cuBLAS "
                         "projections of every layer (seeded random weights) between the ops and "
                         "report it under \"with_gemms\"")
    ap.add_argument("--no-pdl", action="store_true",
                    help="launch without programmatic dependent launch (A/B)")
    ap.add_argument("--no-tick", dest="tick", action="store_false",
                    help="leave the LTR scheduling tick (a8, a10) out of the step")
    ap.add_argument("--live-requests", type=int, default=10000,
                    help="live requests ordered per tick (synthetic-sharegpt.sh submits 10000)")
    ap.add_argument("--tick-prompt-len", type=int, default=245)
    ap.add_argument("--no-extras", action="store_true",
                    help="only the main line: skip ref_cuda / with_gemms / models / trace")
    ap.add_argument("--trace", default="both", choices=["both", "ours", "off"],
                    help="N=1: engine-level trace arms to run (tools/engine_trace.py)")
    ap.add_argument("--trace-model", default="llama-2-7b")
    ap.add_argument("--trace-prompts", type=int, default=10000)
    ap.add_argument("--trace-stop", type=int, default=1000)
    ap.add_argument("--trace-poisson-rate", type=float, default=32.0,
                    help="N=1: second engine-level trace with Poisson arrivals at this many "
                         "requests/s (bench-lmsys.sh / benchmark_serving_real.py:159-176); 0 = off")
    ap.add_argument("--trace-poisson-time", type=float, default=10.0,
                    help="seconds of arrivals (the reference scripts: --request-time 60)")
    ap.add_argument("--trace-poisson-max-output", type=int, default=192,
                    help="output lengths of the trace clipped to this, so that the leg drains "
                         "within the bench's time budget")
    ap.add_argument("--extras-budget-s", type=float, default=540.0,
                    help="stop starting secondary measurements once the run is this old")
    args = ap.parse_args()
    t_run0 = time.time()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from hotpath import SHAPES, ModelShape

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    full_shape = SHAPES[args.model]
    # the reference CPU kernels have no fp16: the benchmark computes in bf16 on both arms
    bench_shape = ModelShape(**{**full_shape.__dict__})
    if bench_shape.dtype == torch.float16:
        bench_shape.dtype = torch.bfloat16
    if args.layers:
        bench_shape.layers = args.layers
    ctx = trace_contexts(args.batch)
    workload = (f"{args.model} decode hot path (a1-a7: rms_norm/fused_add_rms_norm, rotary, "
                f"reshape_and_cache, paged_attention, silu_and_mul x {bench_shape.layers} layers; "
                f"GEMMs excluded), B={args.batch} running sequences of the synthetic ShareGPT "
                f"trace (contexts mean {ctx.mean():.0f}, max {ctx.max()}), block 16"
                + (f"; + LTR tick per step (a8 rank of {args.live_requests} live requests, a10 "
                   f"OPT-125M predictor forward of one {args.tick_prompt_len}-token prompt)"
                   if args.tick else ""))
    config = {"workload": workload, "batch": args.batch, "layers": bench_shape.layers,
              "sum_context_tokens": int(ctx.sum()), "parallelism": f"tp{world}",
              "tp_exchange": (args.tp_exchange if world > 1 else None),
              "l2": "inputs larger than L2 (each layer reads its own KV pool)"}

    # ---------------- reference arm: CPU kernels on host cores, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        tps, ms, desc, cores, kind = cpu_reference_run(args, bench_shape, ctx, args.steps,
                                                       args.warmup, args.cpu_sample_layers,
                                                       args.cpu_sample_batch)
        print(file=out, flush=True, *[json.dumps({
            "impl": "reference", "metric": "output_tokens_per_s", "value": tps,
            "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": kind,
                             "sample": desc},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0})])
        return

    # ---------------- our arm ----------------
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    import torch.distributed as dist
    from vllm_ltr_b200._lib import launch_count, tune
    from hotpath import DecodeHotPath
    if args.no_pdl:
        tune("pdl", 0)
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    shape = bench_shape.shard(world)
    bs = shape.block_size
    need = int(((ctx + bs - 1) // bs).sum())
    num_blocks = need + 64
    hp = DecodeHotPath(shape, num_blocks, args.batch, dev, seed=rank)
    host = build_inputs(ctx, bs, num_blocks, 0)
    max_ctx = int(ctx.max())
    g = torch.Generator().manual_seed(0)
    host["hidden"] = (torch.randn(args.batch, shape.hidden, generator=g) * 0.5).to(shape.dtype)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    out_host = torch.empty(args.batch, shape.hidden, dtype=shape.dtype).pin_memory()

    all_reduce = None
    tp_parity = None
    if world > 1:
        if args.tp_exchange == "fused":
            from vllm_ltr_b200.parallel import FusedTPComm
            hp.attach_fused_tp(FusedTPComm(args.batch, shape.hidden, shape.dtype, dev))
            # parity gate before any timing: a wrong reduction must not produce a number
            tp_parity = tp_parity_check(hp.comm, dev, rank, world, args.batch, shape.hidden,
                                        shape.dtype)
        else:
            from vllm_ltr_b200.parallel import tensor_model_parallel_all_reduce
            all_reduce = tensor_model_parallel_all_reduce

    # the scheduler (and with it the tick) lives on the driver rank only
    tick = GpuTick(dev, args.live_requests, args.tick_prompt_len) if (args.tick and rank == 0) else None

    # Order inside a step: the model stream gets its work first, then the scheduler's tick for the
    # NEXT step (score the prompt that arrived meanwhile, re-rank the queue) runs on its own
    # stream while the model executes, and the host waits for both.
    def step_resident():
        out = hp.step(devin["hidden"], devin["positions"], devin["slot_mapping"],
                      devin["block_tables"], devin["context_lens"], max_ctx, all_reduce)
        if tick:
            tick()
        return out

    # N=1: the step is recorded once into a CUDA graph over static device inputs (the launch
    # mode a latency-sensitive caller uses; PDL edges between the kernels are kept by the graph).
    # `value` replays it with the inputs resident; the e2e leg copies the host buffers into the
    # static inputs, replays, and reads the result back.  N>1 with the fused exchange replays
    # too (every rank in lockstep; the exchange kernels count their barrier epochs on the device);
    # with --tp-exchange nccl the step is issued kernel by kernel.
    use_graph = args.launch == "graph" and (world == 1 or args.tp_exchange == "fused")
    graph = graph_out = None
    ms_no_tick = None
    launches_per_graph = 0
    if use_graph:
        n_before = launch_count()
        graph, graph_out = hp.capture(devin["hidden"], devin["positions"], devin["slot_mapping"],
                                      devin["block_tables"], devin["context_lens"], max_ctx)
        launches_per_graph = (launch_count() - n_before) // 2     # warm-up step + recorded step

    def step_graph():
        graph.replay()
        if tick:
            tick()
        return graph_out

    def step_e2e():
        if use_graph:
            for k, v in pinned.items():
                devin[k].copy_(v, non_blocking=True)
            graph.replay()
            out = graph_out
        else:
            d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            out = hp.step(d["hidden"], d["positions"], d["slot_mapping"], d["block_tables"],
                          d["context_lens"], max_ctx, all_reduce)
        out_host.copy_(out, non_blocking=True)
        if tick:
            tick()
        torch.cuda.current_stream().synchronize()      # the engine samples from this result

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()

    # timed region 1 (eager, inputs resident): every paged_attention call is bracketed by CUDA
    # events on its stream -> the roofline numbers; at N>1 this region is also `value`
    hp.attn_events = []
    n0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.time()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    launches = launch_count() - n0
    ms_eager = ms = e0.elapsed_time(e1)
    attn_ms = [a.elapsed_time(b) for a, b in hp.attn_events]
    hp.attn_events = None

    # attention alone: the step's paged_attention launches (one per layer, each over its own KV
    # pool, so nothing is found in L2) back to back between ONE pair of events - the kernel's
    # in-stream launch duration without the event records that bracket it above (those cost a
    # launch gap and the overlap of the kernel's prologue with its predecessor)
    for _ in range(2):
        hp.attention_only(devin["block_tables"], devin["context_lens"], max_ctx, args.batch)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(args.steps):
        hp.attention_only(devin["block_tables"], devin["context_lens"], max_ctx, args.batch)
    a1.record()
    barrier()
    attn_b2b_ms = a0.elapsed_time(a1) / (args.steps * hp.layers)

    # timed region 1g (N=1, inputs resident): the same K steps replayed from the CUDA graph
    if use_graph:
        for _ in range(2):
            step_graph()
        barrier()
        n0 = launch_count()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            step_graph()
        g1.record()
        barrier()
        launches = launch_count() - n0 + launches_per_graph * args.steps
        ms = g0.elapsed_time(g1)
        # the same replays without the tick on the side stream (what `ref_cuda` is compared with)
        for _ in range(2):
            graph.replay()
        barrier()
        g0.record()
        for _ in range(args.steps):
            graph.replay()
        g1.record()
        barrier()
        ms_no_tick = g0.elapsed_time(g1) / args.steps
    t_end = time.time()

    # host cost of issuing the decode ops of one step (no device wait, tick excluded): must
    # stay below the device time or the step becomes launch-bound
    barrier()
    h0 = time.perf_counter()
    for _ in range(3):
        hp.step(devin["hidden"], devin["positions"], devin["slot_mapping"],
                devin["block_tables"], devin["context_lens"], max_ctx, all_reduce)
    host_ms = (time.perf_counter() - h0) / 3 * 1e3
    barrier()
    tick_ms = None
    if tick:
        h0 = time.perf_counter()
        for _ in range(3):
            tick()
        tick_ms = (time.perf_counter() - h0) / 3 * 1e3      # GPU otherwise idle: tick latency

    # timed region 2: host buffers in, result out, every step
    for _ in range(2):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - w0
    clk = clocks.summary(t_start, t_end)
    clocks2 = clocks          # keeps sampling through the secondary arms; stopped at the end

    if world > 1:
        t = torch.tensor([ms, e2e_s, attn_b2b_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, attn_b2b_ms = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = ms / args.steps
    tokens_per_s = args.batch / (ms_per_step * 1e-3)
    e2e_tps = args.batch * args.steps / e2e_s
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = out_host.numel() * out_host.element_size()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    abytes = attention_bytes(ctx, shape, bs)
    attn_mean_ms = float(np.mean(attn_ms)) if attn_ms else float("nan")
    achieved = abytes / (attn_b2b_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "attention_traffic.json"))).get(
            f"{args.model}-b{args.batch}-tp{world}")
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "paged_attention (v1 entry; one launch: partitions of "
                "long contexts are merged inside the kernel)", "achieved": achieved, "peak": peak,
                "peak_source": peak_src,
                "peak_note": "the measured peak is a COPY (half reads, half writes); a read-only "
                             "stream like this kernel can run a few per cent above it",
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": abytes,
                "launch_ms_mean": attn_b2b_ms,
                "timed_in": f"{args.steps} x {shape.layers} consecutive launches (one per layer, "
                            "each over its own KV pool) between one CUDA event pair",
                "launches_timed": args.steps * shape.layers,
                "launch_ms_bracketed": attn_mean_ms,
                "bracketed_note": "each launch between its own event pair inside the eager step: "
                                  "includes the launch gap the event records introduce",
                "attention_share_of_step": attn_b2b_ms * shape.layers / ms_per_step,
                "eager_ms_per_step": ms_eager / args.steps,
                "step_achieved_gbs": step_bytes(ctx, shape, bs, shape.layers) / (ms_per_step * 1e-3) / 1e9}

    line = {
        "metric": "output_tokens_per_s", "value": tokens_per_s, "unit": "tokens/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": config, "clocks": clk,
        "e2e": {"value": e2e_tps, "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                "launch": ("cuda graph replay (%d kernels per step)" % launches_per_graph
                           if use_graph else "eager")},
        "launch": "cuda graph replay" if use_graph else "eager",
        "pdl": not args.no_pdl,
        "gpu_launches": int(launches), "roofline": roofline,
        "host_issue_ms_per_step": host_ms, "tick_ms_alone": tick_ms,
        "ms_per_step_no_tick": ms_no_tick,
        "metric_note": "`value` is the step rate of the decode HOT PATH (a1-a7 x layers + LTR "
                       "tick; the cuBLAS GEMMs, embedding, lm_head and sampler are not executed): "
                       "batch / time of that kernel chain.  `with_gemms` is the same step with the "
                       "projections in it, and `trace` is the reference's own metric (output "
                       "tokens/s and p50 request latency of the whole engine on the trace).",
    }
    if world > 1:
        line["tp_parity"] = tp_parity
        if args.tp_exchange != "fused":
            line["tp_parity"] = "n/a (NCCL all_reduce)"
    extras = world == 1 and not args.no_extras

    def budget_left():
        return args.extras_budget_s - (time.time() - t_run0)

    def emit():
        if rank == 0:
            print(json.dumps(line), file=out, flush=True)

    if rank == 0:
        # a driver-side timeout must not lose the measurements already taken
        import signal

        def on_term(signum, frame):
            line["terminated_early"] = f"signal {signum} after {time.time() - t_run0:.0f} s"
            emit()
            os._exit(0)

        signal.signal(signal.SIGTERM, on_term)
    ref_cuda = load_ref_cuda() if extras else None
    if extras and ref_cuda is not None:
        # the reference's own CUDA kernels (sm_100a build) on the same step, pools and inputs
        rhp = DecodeHotPath(shape, num_blocks, args.batch, dev, seed=rank, C=ref_cuda)
        for _ in range(3):
            rhp.step(devin["hidden"], devin["positions"], devin["slot_mapping"],
                     devin["block_tables"], devin["context_lens"], max_ctx)
        barrier()
        r0 = time.time()
        rms = time_step_graph(rhp, devin, max_ctx, args.steps)
        r_attn = time_attention_b2b(rhp, devin, max_ctx, args.batch, args.steps)
        r1 = time.time()
        line["ref_cuda"] = {
            "what": "the same decode step (same call order, shapes, inputs, CUDA-graph replay, no "
                    "tick) through the reference's CUDA kernels recompiled unmodified for sm_100a "
                    "(oracle/_ref/ref_cuda_C)",
            "ms_per_step": rms, "tokens_per_s": args.batch / (rms * 1e-3),
            "attention_launch_ms": r_attn,
            "attention_roofline_frac": abytes / (r_attn * 1e-3) / 1e9 / peak,
            "ours_ms_per_step_no_tick": ms_no_tick,
            "speedup_ours_vs_ref_cuda": (rms / ms_no_tick) if ms_no_tick else None,
            "attention_speedup": r_attn / attn_b2b_ms,
            "clocks": clocks2.summary(r0, r1)}
        del rhp
        torch.cuda.empty_cache()
    elif extras:
        line["ref_cuda"] = {"unavailable": "oracle/_ref/ref_cuda_C not built"}
    if (args.with_gemms or extras) and world == 1:
        # context, not the metric: the same step with the dense projections (library code) in it
        hp.enable_gemms()
        g2, out2 = hp.capture(devin["hidden"], devin["positions"], devin["slot_mapping"],
                              devin["block_tables"], devin["context_lens"], max_ctx)
        for _ in range(3):
            g2.replay()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(args.steps):
            g2.replay()
        w1.record()
        barrier()
        gms = w0.elapsed_time(w1) / args.steps
        line["with_gemms"] = {
            "ms_per_step": gms, "tokens_per_s": args.batch / (gms * 1e-3),
            "what": "the same decode step with qkv/o/gate_up/down projections of every layer "
                    "(torch.matmul -> cuBLAS, bf16, seeded random weights, one set per layer) "
                    "between the ops, replayed from a CUDA graph; no embedding, lm_head, sampler "
                    "or scheduler tick", "finite": bool(torch.isfinite(out2.float()).all())}
    # (N=1 only: under torchrun the other ranks would wait in the next collective, and torchrun
    # pins OMP_NUM_THREADS=1, which is not the host the baseline is quoted on)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tps, cms, desc, cores, kind = cpu_reference_run(args, bench_shape, ctx, 2, 1,
                                                        args.cpu_sample_layers,
                                                        args.cpu_sample_batch)
        line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": cores, "kind": kind,
                                "sample": desc, "ms_per_step": cms}
    # ---- N=8: the north-star multi-GPU config (Llama-2-70B, TP8) as a second record ----------
    if world == 8 and args.model != "llama-2-70b" and not args.no_extras and use_graph:
        line["config_70b"] = measure_70b(args, dev, rank, world, peak)
    # ---- N=1: the other single-GPU configs, then the engine-level trace ----------------------
    if extras:
        del hp, graph, graph_out, tick
        hp = graph = graph_out = tick = None
        torch.cuda.empty_cache()
        line["models"] = {}
        for name in ("llama-3-8b", "llama-3-8b-fp8kv"):
            if name == args.model or budget_left() < 240:
                continue
            line["models"][name] = step_pair(name, args.batch, args.steps, dev, ref_cuda, peak,
                                             clocks2, layers=args.layers)
        torch.cuda.empty_cache()
        if args.trace != "off":
            tr = {"workload": f"{args.trace_model} (dummy weights re-drawn N(0,0.02), fp16) + "
                              f"OPT-125M predictor, {args.trace_prompts} requests at t=0 of the "
                              "synthetic ShareGPT trace (BASELINE.md section 4), schedule "
                              f"\"opt-xxx-synthetic((-1,-1,{args.trace_stop}),)\", "
                              "--enable-chunked-prefill --enforce-eager, 768-token / 256-seq "
                              "budget, block 16, greedy, ignore_eos; ordering scores = the "
                              "trace's -log(len)+noise (tau 0.52) in both arms, the predictor "
                              "still runs and is timed; unmodified reference engine "
                              "(oracle/_ref/vllm_py) through LLMEngine.step()",
                  "metric": "output tokens/s and p50 request latency "
                            "(benchmark_throughput_original.py:436-442)"}
            arms = [("ours", "b200")] + ([("ref_cuda", "ref_cuda")] if args.trace == "both" else [])
            for label, arm in arms:
                left = budget_left()
                if left < 90:
                    tr[label] = {"skipped": f"time budget ({args.extras_budget_s:.0f} s) spent"}
                    continue
                tr[label] = run_trace_arm(arm, args, clocks2, min(left, 420.0))
            if "tokens_per_s" in tr.get("ours", {}) and "tokens_per_s" in tr.get("ref_cuda", {}):
                tr["tokens_per_s_ratio"] = tr["ours"]["tokens_per_s"] / tr["ref_cuda"]["tokens_per_s"]
                tr["p50_latency_ratio"] = tr["ours"]["p50_latency_s"] / tr["ref_cuda"]["p50_latency_s"]
            line["trace"] = tr
            if args.trace_poisson_rate > 0:
                tp = {"workload": f"{args.trace_model} + OPT-125M predictor as in `trace`; "
                                  f"requests of the synthetic ShareGPT trace arriving with "
                                  f"Gamma(cv=1) gaps at {args.trace_poisson_rate:g} requests/s for "
                                  f"{args.trace_poisson_time:g} s (benchmark_serving_real.py:159-176)"
                                  f", outputs clipped to {args.trace_poisson_max_output} tokens, "
                                  "schedule \"opt-xxx\", every request run to its end; a request "
                                  "enters the engine at the first step boundary after its arrival",
                      "metric": "output tokens/s and p50 request latency from arrival"}
                for label, arm in arms:
                    left = budget_left()
                    if left < 0.5 * args.extras_budget_s:     # last leg: only in the first half
                        tp[label] = {"skipped": "more than half of the time budget "
                                                f"({args.extras_budget_s:.0f} s) spent"}
                        continue
                    tp[label] = run_trace_arm(arm, args, clocks2, min(left, 150.0), poisson=True)
                if "p50_latency_s" in tp.get("ours", {}) and "p50_latency_s" in tp.get("ref_cuda", {}):
                    tp["tokens_per_s_ratio"] = (tp["ours"]["tokens_per_s"]
                                                / tp["ref_cuda"]["tokens_per_s"])
                    tp["p50_latency_ratio"] = (tp["ours"]["p50_latency_s"]
                                               / tp["ref_cuda"]["p50_latency_s"])
                line["trace_poisson"] = tp
    clocks2.stop()
    line["bench_wall_s"] = time.time() - t_run0
    emit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
