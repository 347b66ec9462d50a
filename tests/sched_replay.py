"""Replay harness for the LTR scheduling tick: drives a reference ``Scheduler`` (from
/root/reference/vllm/core/scheduler.py) through a seeded request population and a simulated
engine (the bookkeeping of LLMEngine._process_model_outputs, llm_engine.py:471-510, without a
model), recording every tick's decision.  Used twice per scenario - once with the reference's own
methods, once with vllm_ltr_b200.ltr_scheduler installed - and the two recordings must be
identical."""
import os
import sys

import numpy as np

# the reference's Python package: the checkout in the build container, or the unmodified copy
# `make -C oracle vllm_py` ships to the GPU box (oracle/_ref/vllm_py, git-ignored)
_SHIPPED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "oracle", "_ref", "vllm_py")
REF = "/root/reference" if os.path.isdir("/root/reference/vllm/core") else _SHIPPED


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF, "vllm", "core"))


def import_reference():
    """Make `import vllm` resolve to the reference checkout (not the vllm wheel in the image)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "vllm" or k.startswith("vllm.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(REF) and not k.startswith("vllm._C"):
            del sys.modules[k]
    import vllm  # noqa: F401
    assert vllm.__file__.startswith(REF), vllm.__file__


class StubScorer:
    """scheduler.aux_model stand-in: fp16-quantised scores keyed by request id (many exact ties,
    like `logits[:,0].tolist()` of the predictor, opt.py:399-409)."""

    def __init__(self, scores):
        self.scores = scores
        self.calls = []

    def obtain_aux_scores(self, seq_groups):
        self.calls.append([g.request_id for g in seq_groups])
        out = []
        for g in seq_groups:
            assert g.need_aux_model_score()
            g.set_aux_model_score(self.scores[g.request_id])
            out.append(self.scores[g.request_id])
        return out


class PythonRanker:
    """Host-logic test double for the CUDA ranker: the reference's own expression."""

    def order(self, scores, pri=None, negate=True):
        n = len(scores)
        sgn = -1.0 if negate else 1.0
        if pri is None:
            idx = sorted(range(n), key=lambda i: sgn * scores[i])
        else:
            idx = sorted(range(n), key=lambda i: (pri[i], sgn * scores[i]))
        return np.asarray(idx, dtype=np.int32)


import time as _time


class NumpyRanker:
    """Stable order of (pri, -score): what the CUDA rank kernel returns (tests/test_gpu_ltr.py)."""
    seconds = 0.0      # accumulated time spent ranking (tools/sched_tick_bench.py reports it)

    def order(self, scores, pri=None, negate=True):
        t0 = _time.perf_counter()
        key = -np.asarray(scores, dtype=np.float64) if negate else np.asarray(scores, dtype=np.float64)
        if pri is None:
            idx = np.argsort(key, kind="stable")
        else:
            idx = np.lexsort((key, np.asarray(pri)))       # last key is the primary one; stable
        NumpyRanker.seconds += _time.perf_counter() - t0
        return idx.astype(np.int32)


def make_population(n, seed, max_prompt=96, max_out=48):
    rng = np.random.default_rng(seed)
    prompt = np.clip(np.rint(np.exp(rng.normal(3.0, 0.8, n))), 2, max_prompt).astype(int)
    out = np.clip(np.rint(np.exp(rng.normal(2.5, 0.9, n))), 1, max_out).astype(int)
    score = (-np.log(out) + rng.normal(0, 0.7, n)).astype(np.float16).astype(np.float64)
    arrival = np.sort(rng.integers(0, max(1, n // 8), n))      # tick at which each arrives
    return prompt, out, score, arrival


def run(n, seed, schedule_type, num_gpu_blocks, num_cpu_blocks, block_size=4, max_tokens=64,
        max_seqs=16, use_plugin=False, ranker=None, max_ticks=5000, aborts=None,
        max_prompt=96, on_tick=None):
    """Returns the list of per-tick records.  `aborts` maps a tick number to the request ids
    aborted (Scheduler.abort_seq_group, scheduler.py:379-413) before that tick's schedule();
    `on_tick(sched, groups, tick)` runs after every simulated engine step."""
    import_reference()
    from vllm import SamplingParams
    from vllm.config import CacheConfig, SchedulerConfig
    from vllm.core.scheduler import Scheduler
    from vllm.sequence import Logprob, Sequence, SequenceGroup, SequenceStatus

    prompt, out_len, score, arrival = make_population(n, seed, max_prompt=max_prompt)
    sc = SchedulerConfig(max_tokens, max_seqs, 4096, enable_chunked_prefill=True,
                         schedule_type=schedule_type)
    cc = CacheConfig(block_size, 1.0, 1, "auto")
    cc.num_gpu_blocks, cc.num_cpu_blocks = num_gpu_blocks, num_cpu_blocks
    sched = Scheduler(sc, cc, None)
    sched.aux_model = StubScorer({str(i): float(score[i]) for i in range(n)})
    if use_plugin:
        from vllm_ltr_b200 import ltr_scheduler
        ltr_scheduler.install(sched, ranker=ranker)
    groups = []
    for i in range(n):
        toks = list(range(int(prompt[i])))
        seq = Sequence(i, " ".join(map(str, toks)), toks, block_size)
        sp = SamplingParams(ignore_eos=True, max_tokens=int(out_len[i]))
        groups.append(SequenceGroup(str(i), [seq], sp, 0.0))
    records = []
    nxt = 0
    tick = 0
    while tick < max_ticks:
        while nxt < n and arrival[nxt] <= tick:
            sched.add_seq_group(groups[nxt])
            nxt += 1
        if aborts and tick in aborts:
            sched.abort_seq_group([str(i) for i in aborts[tick]])
        if not sched.has_unfinished_seqs():
            if nxt >= n:
                break
            tick += 1
            continue
        metas, outs = sched.schedule()
        rec = {
            "sched": [(s.seq_group.request_id, s.token_chunk_size, m.is_prompt)
                      for s, m in zip(outs.scheduled_seq_groups, metas)],
            # physical block NUMBERS are not reproducible even between two runs of the reference:
            # BlockSpaceManagerV1.free iterates `set(block_table)` (block_manager_v1.py:560-564),
            # i.e. in object-address order, so the free list is permuted run to run.  Sizes are.
            "block_tables": [sorted((k, len(v)) for k, v in m.block_tables.items())
                             for m in metas],
            # everything else schedule() puts into the SequenceGroupMetadata (scheduler.py:1661-1678)
            "meta": [(m.request_id, m.token_chunk_size,
                      sorted((k, d.get_len(), d is s.seq_group.seqs_dict[k].data)
                             for k, d in m.seq_data.items()),
                      list(m.computed_block_nums), m.need_score, m.multi_modal_data is None,
                      m.state is s.seq_group.state, m.sampling_params is s.seq_group.sampling_params,
                      m.lora_request is None,
                      s.seq_group.metrics.first_scheduled_time is not None,
                      s.seq_group.metrics.time_in_queue is not None)
                     for s, m in zip(outs.scheduled_seq_groups, metas)],
            "n_prefill": outs.num_prefill_groups,
            "n_tokens": outs.num_batched_tokens,
            "swap_in": len(outs.blocks_to_swap_in),
            "swap_out": len(outs.blocks_to_swap_out),
            "copy": sorted(len(v) for v in outs.blocks_to_copy.values()),
            "free_gpu": sched.block_manager.get_num_free_gpu_blocks(),
            "free_cpu": sched.block_manager.get_num_free_cpu_blocks(),
            "waiting": [g.request_id for g in sched.waiting],
            "running": [g.request_id for g in sched.running],
            "swapped": [g.request_id for g in sched.swapped],
            "state": [(g.request_id, g.idle, g.runs, g.pri, g.running_info["swap_out"],
                       g.running_info["swap_blocks"], hasattr(g, "num_new_tokens"))
                      for g in list(sched.swapped) + list(sched.running) + list(sched.waiting)],
        }
        records.append(rec)
        # simulated engine step (llm_engine.py:489-499): computed-token bookkeeping, one sampled
        # token for every group whose prompt is fully computed, stop at max_tokens
        for s in outs.scheduled_seq_groups:
            g = s.seq_group
            g.update_num_computed_tokens(s.token_chunk_size)
            if g.get_num_uncomputed_tokens() == 0:
                for seq in g.get_seqs(status=SequenceStatus.RUNNING):
                    seq.append_token_id(7, {7: Logprob(0.0)})
                    if seq.get_output_len() >= g.sampling_params.max_tokens:
                        seq.status = SequenceStatus.FINISHED_LENGTH_CAPPED
                        sched.free_seq(seq)
        sched.free_finished_seq_groups()
        if on_tick is not None:
            on_tick(sched, groups, tick)
        tick += 1
    return records
