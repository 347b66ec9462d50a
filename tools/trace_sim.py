#!/usr/bin/env python
"""Trace replay of the reference's headline metric (output tokens/s and p50 request latency,
benchmarks/benchmark_throughput_original.py:436-442) for the LTR policy - a MODEL, not a
measurement of an engine: the host code of every step is executed and timed for real on this
machine's CPU, the GPU time of the step is a two-parameter model fitted to the kernel
measurements in profiles/.

Per simulated step:
  host   Scheduler.schedule() of the reference (/root/reference/vllm/core/scheduler.py), with its
         own LTR methods or with vllm_ltr_b200.ltr_scheduler installed   - executed, wall time
         ModelRunner._prepare_decode (reference) or vllm_ltr_b200.decode_inputs.prepare_decode on
         the step's decode requests                                         - executed, wall time
         the sampler's host-side tensor preparation: the reference's
         SamplingTensors.from_sampling_metadata (sampling_metadata.py:79-330, pads every
         sequence's prompt and output tokens into [batch, max_len] matrices even when no penalty
         is requested) or the two per-row lists B200Sampler.forward builds    - executed, wall time
  gpu    t = base + per_ctx_token * (sum of the contexts of the scheduled sequences)
         ours:      base 4.6 ms, 25 ns/token   (Llama-3-8B: 3.2 ms cuBLAS projections + 1.4 ms small
                    kernels, attention 82 us x 32 layers per 103 k context tokens;
                    profiles/r01_bench_n1_llama3-8b*.json)
         reference: base 4.3 ms, 97 ns/token   (its CUDA kernels recompiled for sm_100a: 313 us x 32
                    per 103 k context tokens; profiles/r01_step_vs_reference_cuda.txt)
Scores: -log(output_len) + noise tuned to Kendall tau ~ 0.5, fp16; the
predictor forward itself is not in the loop (stub scorer on both arms).  Requests: the synthetic
ShareGPT-shaped trace, all present at t = 0 (burst-lmsys.sh / synthetic-sharegpt.sh).

    python tools/trace_sim.py [--n 2000] [--stop-after 0] > profiles/r01_trace_sim.txt
"""
import argparse
import logging
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import sched_replay as R  # noqa: E402

R.import_reference()
logging.disable(logging.CRITICAL)
import numpy as np  # noqa: E402
from vllm import SamplingParams  # noqa: E402
from vllm.config import CacheConfig, SchedulerConfig  # noqa: E402
from vllm.core.scheduler import Scheduler  # noqa: E402
from vllm.sequence import Logprob, Sequence, SequenceGroup, SequenceStatus  # noqa: E402
from vllm.worker.model_runner import ModelRunner  # noqa: E402
import torch  # noqa: E402
import vllm.model_executor.sampling_metadata as smd  # noqa: E402
from vllm.model_executor.sampling_metadata import SamplingMetadata, SamplingTensors  # noqa: E402
from vllm.sampling_params import SamplingType  # noqa: E402

smd.is_pin_memory_available = lambda: False      # no CUDA driver in the build container

from vllm_ltr_b200 import decode_inputs, ltr_scheduler  # noqa: E402

NumpyRanker = R.NumpyRanker     # stand-in for the CUDA rank kernel (no GPU in this container)

GPU_MODEL = {"ours": (4.6e-3, 25e-9), "reference": (4.3e-3, 97e-9)}


def population(n, seed=0):
    rng = np.random.default_rng(seed)
    prompt = np.clip(np.rint(np.exp(rng.normal(5.0, 1.0, n))), 4, 2048).astype(int)
    out = np.clip(np.rint(np.exp(rng.normal(5.3, 1.1, n))), 4, 8192 - prompt).astype(int)
    score = (-np.log(out) + rng.normal(0, 1.45, n)).astype(np.float16).astype(np.float64)
    return prompt, out, score


def run(arm, n, stop_after, gpu_blocks, sampler_host=True):
    prompt, out_len, score = population(n)
    sc = SchedulerConfig(768, 256, 8192, enable_chunked_prefill=True, schedule_type="opt-xxx")
    cc = CacheConfig(16, 1.0, 1, "auto")
    cc.num_gpu_blocks, cc.num_cpu_blocks = gpu_blocks, 4 * gpu_blocks
    sched = Scheduler(sc, cc, None)
    sched.aux_model = R.StubScorer({str(i): float(score[i]) for i in range(n)})
    fake_runner = types.SimpleNamespace(
        sliding_window=None, block_size=16, device="cpu",
        model_config=types.SimpleNamespace(enforce_eager=True), max_context_len_to_capture=8192,
        attn_backend=types.SimpleNamespace(make_metadata=lambda **kw: None))
    if arm == "ours":
        ltr_scheduler.install(sched, ranker=NumpyRanker())
        prepare = decode_inputs.prepare_decode
    else:
        prepare = ModelRunner._prepare_decode
    for i in range(n):
        toks = list(range(int(prompt[i])))
        sched.add_seq_group(SequenceGroup(str(i), [Sequence(i, "", toks, 16)],
                                          SamplingParams(ignore_eos=True, max_tokens=int(out_len[i])),
                                          0.0))
    base, per_tok = GPU_MODEL[arm]
    now, host_s, gpu_s, steps, generated = 0.0, 0.0, 0.0, 0, 0
    finish = {}
    while sched.has_unfinished_seqs():
        t0 = time.perf_counter()
        metas, outs = sched.schedule()
        decode = [m for m in metas if not m.is_prompt]
        if decode:
            prepare(fake_runner, decode)
        if sampler_host:
            if arm == "ours":      # what B200Sampler.forward does on the host before its kernel
                temps, rnd = [], []
                for m in metas:
                    sp = m.sampling_params
                    k = len(m.seq_data)
                    temps += [sp.temperature if sp.temperature >= 1e-5 else 1.0] * k
                    rnd += [int(sp.sampling_type == SamplingType.RANDOM)] * k
            else:
                seq_data = {}
                for m in metas:
                    seq_data.update(m.seq_data)
                md = SamplingMetadata([(list(m.seq_data.keys()), m.sampling_params) for m in metas],
                                      seq_data, [], torch.empty(0, dtype=torch.long), None)
                SamplingTensors.from_sampling_metadata(md, 32000, torch.device("cpu"), torch.float16)
        h = time.perf_counter() - t0
        ctx = 0
        for s_ in outs.scheduled_seq_groups:
            g = s_.seq_group
            ctx += g.get_seqs()[0].data.get_num_computed_tokens() + s_.token_chunk_size
        gpu = base + per_tok * ctx
        now += h + gpu
        host_s += h
        gpu_s += gpu
        steps += 1
        for s_ in outs.scheduled_seq_groups:
            g = s_.seq_group
            g.update_num_computed_tokens(s_.token_chunk_size)
            if g.get_num_uncomputed_tokens() == 0:
                for seq in g.get_seqs(status=SequenceStatus.RUNNING):
                    seq.append_token_id(7, {7: Logprob(0.0)})
                    generated += 1
                    if seq.get_output_len() >= g.sampling_params.max_tokens:
                        seq.status = SequenceStatus.FINISHED_LENGTH_CAPPED
                        sched.free_seq(seq)
                        finish[g.request_id] = now
        sched.free_finished_seq_groups()
        if stop_after and len(finish) >= stop_after:
            break
    lat = np.array(sorted(finish.values()))
    return dict(arm=arm, steps=steps, completed=len(finish), generated=generated, elapsed=now,
                tokens_per_s=generated / now, p50=float(np.percentile(lat, 50)),
                p90=float(np.percentile(lat, 90)), host_ms=host_s / steps * 1e3,
                gpu_ms=gpu_s / steps * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--stop-after", type=int, default=0, help="stop after this many completions")
    ap.add_argument("--no-sampler-host", dest="sampler_host", action="store_false",
                    help="leave the sampler's host-side tensor preparation out of the step")
    ap.add_argument("--gpu-blocks", type=int, default=60000,
                    help="KV pool in 16-token blocks (Llama-3-8B bf16 on 180 GB: ~60 k)")
    args = ap.parse_args()
    print(f"# simulated LTR run: {args.n} requests at t=0, stop after "
          f"{args.stop_after or 'all'} completions, {args.gpu_blocks} KV blocks; host code executed "
          f"and timed on this CPU, GPU step time modelled (see the header of tools/trace_sim.py)"
          f"{'' if args.sampler_host else '; sampler host work excluded'}")
    for arm in ("reference", "ours"):
        r = run(arm, args.n, args.stop_after, args.gpu_blocks, args.sampler_host)
        print(f"{r['arm']:9s}: {r['completed']} requests, {r['generated']} output tokens in "
              f"{r['elapsed']:.1f} s simulated = {r['tokens_per_s']:.0f} tokens/s; request latency "
              f"p50 {r['p50']:.1f} s, p90 {r['p90']:.1f} s; {r['steps']} steps, per step: host "
              f"{r['host_ms']:.1f} ms (schedule + decode inputs"
              f"{' + sampler tensors' if args.sampler_host else ''}, measured), GPU {r['gpu_ms']:.1f} ms "
              f"(model)", flush=True)


if __name__ == "__main__":
    main()
