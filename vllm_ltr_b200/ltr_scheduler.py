"""The learning-to-rank scheduling tick, as a plugin for the reference's ``Scheduler``.

The reference binds three callables on its scheduler from the ``schedule_type`` string
(vllm/core/scheduler.py:290-331): ``_schedule``, ``_get_ordered_requests`` and
``_update_priority``.  For ``schedule_type`` "opt*" they are ``_general_schedule`` (:1101-1373,
with ``reserve_free_blocks`` :1376-1502) and ``_get_opt_ordered_requests`` (:969-1000).
``install(scheduler)`` re-binds those three names to the functions below.  They produce, tick by
tick, the same ``SchedulerOutputs``, the same ``waiting/running/swapped`` deques and the same
``SequenceGroup.{idle,runs,pri,num_new_*}`` side effects as the reference methods (replay-tested
against ``/root/reference/vllm/core/scheduler.py``), but

* the queue sort runs on the GPU (``vltr_ltr_rank_host``, csrc/ltr_rank.cu) instead of Python's
  ``sorted`` over key tuples, and
* the bookkeeping around it is linear in the number of live requests with a small constant: the
  reference scans deques for membership (``x in deque``, :1139/:1166/:1187), removes the selected
  requests one ``list.remove`` at a time (:1213-1214) and tests every live request against the
  list of scheduled ones (:1358-1365) - about 100-140 ms per tick at 10 000 live requests.  Here
  only running / swapped requests (hundreds) are looked up by identity, the queue a rank came
  from is a range test on the rank permutation, the selected requests are a prefix of the order,
  and the waiting queue is rebuilt with one C-level ``compress``.

Quirks of the reference that are part of its observable behaviour and therefore kept:
* ties in (pri, -score) keep waiting -> running -> swapped deque order (stable sort);
* preemption is always by swap (``PreemptionMode.SWAP``, :1410,1434);
* proactively swapped-in requests are NOT charged to ``final_budget``: :1491-1492 re-add the last
  pinned request's id, which the budget ignores as a duplicate (and raises ``UnboundLocalError``
  when nothing was pinned) - see ``_reserve_free_blocks``.
"""
import os
import time
import types
from collections import deque
from itertools import compress, islice
from operator import itemgetter
from typing import Dict, Iterable, List, Optional

import numpy as np

from ._lib import check, lib

_WAITING, _RUNNING, _SWAPPED = 0, 1, 2


class GpuRanker:
    """Stable order of (pri, -score) on the GPU: the replacement for ``sorted`` at
    scheduler.py:996/998.  ``order()`` returns the permutation as a numpy int32 array."""

    def __init__(self, device: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("vllm_ltr_b200 LTR ranker needs a CUDA device (no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else device
        self._torch = torch
        self._stream = torch._C._cuda_getCurrentRawStream

    def order(self, scores, pri=None, negate: bool = True) -> np.ndarray:
        n = len(scores)
        out = np.empty(n, dtype=np.int32)
        if n == 0:
            return out
        s = np.ascontiguousarray(scores, dtype=np.float64)
        p = None if pri is None else np.ascontiguousarray(pri, dtype=np.int32)
        with self._torch.cuda.device(self.device):
            check(lib.vltr_ltr_rank_host(
                s.ctypes.data, None if p is None else p.ctypes.data, n, 1 if negate else 0,
                out.ctypes.data, self._stream(self.device)))
        return out


def _ref():
    """Types of the engine this plugin drops into (the reference's own module)."""
    import vllm.core.scheduler as S
    return S


# ---------------------------------------------------------------------------------------------
# _get_ordered_requests for schedule_type "opt*"        (scheduler.py:969-1000)
# ---------------------------------------------------------------------------------------------
def get_opt_ordered_requests(self) -> list:
    # need_aux_model_score() is `aux_model_score is None` (vllm/sequence.py:460-461): read the
    # attribute directly.  One pass gathers the waiting queue's scores (the sort needs them
    # anyway); the list of unscored requests is only built on the ticks that have any.
    live = list(self.waiting)
    nw = len(live)
    scores = [r.aux_model_score for r in live]
    if None in scores:
        unscored = [r for r, sc in zip(live, scores) if sc is None]
        timed = int(os.environ.get("OPT_TIME", 0))
        t0 = time.time() if timed else 0.0
        self.aux_model.obtain_aux_scores(unscored)
        if timed:
            print("OPT-TIME: ", time.time() - t0)
        scores = [r.aux_model_score for r in live]
    live.extend(self.running)
    nr = len(live) - nw
    live.extend(self.swapped)
    scores.extend([r.aux_model_score for r in islice(live, nw, None)])
    pri = None
    if self.starv != -1:
        # starvation control: a request idle for `starv` ticks is promoted for `period`
        # scheduled steps, then demoted again
        for r in live:
            if r.idle >= self.starv:
                r.pri, r.idle, r.runs = -1, 0, self.period
            elif r.pri == -1 and r.runs <= 0:
                r.pri = 0
        pri = [r.pri for r in live]
    order = self._vltr_ranker.order(scores, pri, negate=True)
    n = len(live)
    if n >= 2:
        ordered = list(itemgetter(*order.tolist())(live))
    else:
        ordered = list(live)
    # remembered for general_schedule: which queue every rank came from is a range test on the
    # permutation (live = waiting ++ running ++ swapped), no per-request lookups
    self._vltr_tick = (ordered, live, order, nw, nr)
    return ordered


def update_opt_priority(self) -> None:     # scheduler.py:1002-1003
    return None


# ---------------------------------------------------------------------------------------------
# reserve_free_blocks                                    (scheduler.py:1376-1502)
# ---------------------------------------------------------------------------------------------
def _reserve_free_blocks(self, S, blocks_needed: int, pinned: list,
                         evict_candidates: Iterable, swapin_candidates: Iterable,
                         where: Dict[int, int], dropped_running: set, final_budget):
    """Make room for the pinned (selected) requests: swap out from the tail of the rank order,
    then from the tail of the pinned list; otherwise swap in the best-ranked swapped requests.
    `evict_candidates`: the members of the running queue among the unselected requests, LOWEST
    rank first; `swapin_candidates`: the members of the swapped queue among them, best rank first
    (the reference walks all unselected requests and skips the others, :1402-1416 / :1459-1499).
    `where` maps id(request) -> queue at the start of the tick for running / swapped requests
    (waiting ones are absent); requests preempted here are recorded in `dropped_running` (they
    leave self.running).  Returns
    (execute, preempted, swapped_out, blocks_to_swap_out, blocks_to_swap_in)."""
    bm = self.block_manager
    RUNNING, SWAPPED = S.SequenceStatus.RUNNING, S.SequenceStatus.SWAPPED
    to_out: Dict[int, int] = {}
    to_in: Dict[int, int] = {}
    preempted, swapped_out = [], []
    shortfall = blocks_needed - bm.gpu_allocator.get_num_free_blocks() + bm.watermark_blocks
    must_swap_out = shortfall > 0
    execute = list(pinned)

    def evict(req):
        mode = self._preempt(req, to_out, preemption_mode=S.PreemptionMode.SWAP)
        (preempted if mode == S.PreemptionMode.RECOMPUTE else swapped_out).append(req)

    if must_swap_out:
        for req in evict_candidates:                    # lowest rank first
            if shortfall <= 0:
                break
            # only members of the running queue can hold RUNNING sequences (waiting / swapped
            # requests hold WAITING / SWAPPED ones)
            if req.get_seqs(status=RUNNING):
                shortfall -= len(bm._get_physical_blocks(req))
                evict(req)
                if where.get(id(req)) == _RUNNING and id(req) not in dropped_running:
                    dropped_running.add(id(req))
                else:
                    execute.remove(req)                 # reference: list.remove -> ValueError
        while shortfall > 0 and execute:                # still short: give up pinned requests
            req = execute.pop()
            if req.get_seqs(status=RUNNING):
                shortfall -= req.num_seqs(status=RUNNING)
                shortfall -= len(bm._get_physical_blocks(req))
                evict(req)
                if where.get(id(req)) != _RUNNING or id(req) in dropped_running:
                    raise ValueError("deque.remove(x): x not in deque")   # as the reference
                dropped_running.add(id(req))
            elif req.get_seqs(status=SWAPPED):
                shortfall -= len(bm._get_physical_blocks(req)) + req.num_seqs(status=SWAPPED)
            else:
                shortfall -= len(req.get_seqs()[0].logical_token_blocks)
        assert shortfall <= 0

    last_pinned = None
    for req in execute:
        if req.get_seqs(status=SWAPPED):
            self._swap_in(req, to_in)
        final_budget.add_num_batched_tokens(req.request_id, req.num_new_tokens)
        final_budget.add_num_seqs(req.request_id, req.num_new_seqs)
        last_pinned = req

    if not must_swap_out:
        # proactive swap-in, best rank first, while blocks (shortfall is <= 0 here: the spare
        # room) and the final budget allow
        for req in swapin_candidates:
            if not req.get_seqs(status=SWAPPED):
                continue
            new_seqs = req.get_max_num_running_seqs()
            new_tokens = self._get_num_new_tokens(req, SWAPPED, enable_chunking=True,
                                                  budget=final_budget)
            cost = len(bm._get_physical_blocks(req)) + req.num_seqs(status=SWAPPED)
            if (bm.can_swap_in(req) and shortfall + cost < 0 and new_tokens > 0
                    and final_budget.can_schedule(num_new_tokens=new_tokens,
                                                  num_new_seqs=new_seqs)):
                req.num_new_seqs = req.get_max_num_running_seqs()
                req.num_new_tokens = sum(s.get_num_new_tokens() for s in req.get_seqs(status=SWAPPED))
                self._swap_in(req, to_in)
                if last_pinned is None:
                    # scheduler.py:1491 reads the loop variable of an empty loop
                    raise UnboundLocalError(
                        "cannot access local variable 'seq_group' where it is not associated "
                        "with a value")
                # scheduler.py:1491-1492 charge `seq_group` (the last pinned request), whose id
                # the budget has already seen: a no-op, kept as such
                final_budget.add_num_batched_tokens(last_pinned.request_id,
                                                    last_pinned.num_new_tokens)
                final_budget.add_num_seqs(last_pinned.request_id, last_pinned.num_new_seqs)
                execute.append(req)
                shortfall += cost
            else:
                break
    return execute, preempted, swapped_out, to_out, to_in


# ---------------------------------------------------------------------------------------------
# _schedule for the rank-ordered policies                (scheduler.py:1101-1373)
# ---------------------------------------------------------------------------------------------
def general_schedule(self):
    S = _ref()
    WAITING, RUNNING, SWAPPED = (S.SequenceStatus.WAITING, S.SequenceStatus.RUNNING,
                                 S.SequenceStatus.SWAPPED)
    self._update_priority()
    ordered = self._get_ordered_requests()
    live_before = len(self.waiting) + len(self.running) + len(self.swapped)
    cfg = self.scheduler_config
    budget = S.SchedulingBudget(token_budget=cfg.max_num_batched_tokens,
                                max_num_seqs=cfg.max_num_seqs)
    final_budget = S.SchedulingBudget(token_budget=cfg.max_num_batched_tokens,
                                      max_num_seqs=cfg.max_num_seqs)
    # which queue a request sits in.  Running and swapped requests (at most max_num_seqs plus
    # whatever was preempted - hundreds) are looked up by identity; everything else in the order
    # is a waiting request.  When the order came from get_opt_ordered_requests the ranks of the
    # running / swapped requests are also known from the permutation itself (range tests on the
    # indices into waiting ++ running ++ swapped), so no pass over the ~10^4 waiting requests is
    # needed anywhere below.
    tick = getattr(self, "_vltr_tick", None)
    fast = tick is not None and tick[0] is ordered and len(ordered) == live_before
    where: Dict[int, int] = {id(r): _RUNNING for r in self.running}
    for r in self.swapped:
        where[id(r)] = _SWAPPED
    if not fast:
        for r in self.waiting:
            where[id(r)] = _WAITING
    default_tag = _WAITING if fast else None

    # ---- admit in rank order until a budget runs out: the selection is a prefix of the order
    bm = self.block_manager
    blocks_needed = 0
    k = 0
    for req in ordered:
        tag = where.get(id(req), default_tag)
        if tag == _RUNNING:
            new_tokens = self._get_num_new_tokens(req, RUNNING, True, budget)
            if new_tokens == 0:
                assert budget.remaining_token_budget() == 0
                break
            new_seqs = req.get_max_num_running_seqs()
            if not budget.can_schedule(num_new_tokens=new_tokens, num_new_seqs=new_seqs):
                break
            blocks = new_seqs
        elif tag == _SWAPPED:
            new_seqs = req.get_max_num_running_seqs()
            new_tokens = self._get_num_new_tokens(req, SWAPPED, True, budget)
            n_swapped = req.num_seqs(status=SWAPPED)
            if new_tokens == 0 or not budget.can_schedule(num_new_tokens=new_tokens,
                                                          num_new_seqs=new_seqs):
                break
            blocks = len(bm._get_physical_blocks(req)) + n_swapped
        elif tag == _WAITING:
            new_tokens = self._get_num_new_tokens(req, WAITING, True, budget)
            assert new_tokens <= self.prompt_limit, "req exceed prompt limit"
            new_seqs = req.get_max_num_running_seqs()
            if new_tokens == 0 or not budget.can_schedule(num_new_tokens=new_tokens,
                                                          num_new_seqs=new_seqs):
                break
            blocks = len(req.get_seqs()[0].logical_token_blocks)
        else:
            raise AssertionError("seqgroup not in all lists")
        budget.add_num_batched_tokens(req.request_id, new_tokens)
        budget.add_num_seqs(req.request_id, new_seqs)
        req.num_new_tokens, req.num_new_seqs = new_tokens, new_seqs
        blocks_needed += blocks
        k += 1
    selected = ordered[:k]
    if fast:
        _, live, order, nw, nr = tick
        tail = order[k:]
        run_ranks = np.flatnonzero((tail >= nw) & (tail < nw + nr))
        swp_ranks = np.flatnonzero(tail >= nw + nr)
        evict_candidates = (ordered[k + i] for i in run_ranks[::-1].tolist())
        swapin_candidates = (ordered[k + i] for i in swp_ranks.tolist())
    else:
        rest = ordered[k:]
        evict_candidates = (r for r in reversed(rest) if where.get(id(r)) == _RUNNING)
        swapin_candidates = (r for r in rest if where.get(id(r)) == _SWAPPED)

    dropped_running: set = set()      # ids leaving self.running by preemption
    execute, preempted, swapped_out, to_swap_out, to_swap_in = _reserve_free_blocks(
        self, S, blocks_needed, selected, evict_candidates, swapin_candidates, where,
        dropped_running, final_budget)

    # ---- allocate / append slots for what really executes -------------------------------------
    to_copy: Dict[int, List[int]] = {}
    new_prefill, run_prefill, run_decode, swp_prefill, swp_decode = [], [], [], [], []
    taken = {_WAITING: set(), _RUNNING: set(), _SWAPPED: set()}
    SSG = S.ScheduledSequenceGroup
    for req in execute:
        tag = where.get(id(req), default_tag)
        if tag == _RUNNING and id(req) in dropped_running:
            tag = None
        if tag is None or id(req) in taken[tag]:
            raise AssertionError
        taken[tag].add(id(req))
        if tag == _WAITING:
            if bm.can_allocate(req) != S.AllocStatus.OK:
                raise AssertionError("can not append new req")
            self._allocate_and_set_running(req, req.num_new_tokens)
            new_prefill.append(SSG(seq_group=req, token_chunk_size=req.num_new_tokens))
        else:
            if not bm.can_append_slots(req):
                raise AssertionError
            self._append_slots(req, to_copy)
            if req.is_prefill():
                (run_prefill if tag == _RUNNING else swp_prefill).append(
                    SSG(seq_group=req, token_chunk_size=req.num_new_tokens))
            else:
                if tag == _SWAPPED:
                    assert req.num_new_tokens == 1
                (run_decode if tag == _RUNNING else swp_decode).append(
                    SSG(seq_group=req, token_chunk_size=1))
        del req.num_new_tokens
        del req.num_new_seqs

    assert final_budget.num_batched_tokens <= cfg.max_num_batched_tokens
    assert budget.num_curr_seqs <= cfg.max_num_seqs

    # ---- rebuild the three queues in one pass each ----------------------------------------------
    gone_w, gone_r, gone_s = taken[_WAITING], taken[_RUNNING] | dropped_running, taken[_SWAPPED]
    if not gone_w:
        waiting = deque(self.waiting)
    elif fast:
        # the admitted waiting requests sit at known positions of the old queue: order[rank]
        keep = bytearray(b"\x01") * nw
        for rank, req in enumerate(selected):
            if id(req) in gone_w:
                keep[order[rank]] = 0
        waiting = deque(compress(islice(live, nw), keep))
    else:
        waiting = deque(r for r in self.waiting if id(r) not in gone_w)
    waiting.extendleft(preempted)
    running = deque(r for r in self.running if id(r) not in gone_r)
    for group in (new_prefill, run_decode, run_prefill, swp_decode, swp_prefill):
        running.extend(s.seq_group for s in group)
    swapped = deque(r for r in self.swapped if id(r) not in gone_s)
    swapped.extend(swapped_out)
    self.waiting, self.running, self.swapped = waiting, running, swapped
    assert len(waiting) + len(running) + len(swapped) == live_before

    scheduled = new_prefill + run_prefill + swp_prefill + run_decode + swp_decode
    out = S.SchedulerOutputs(
        scheduled_seq_groups=scheduled,
        num_prefill_groups=len(new_prefill) + len(swp_prefill) + len(run_prefill),
        num_batched_tokens=final_budget.num_batched_tokens,
        blocks_to_swap_in=to_swap_in,
        blocks_to_swap_out=to_swap_out,
        blocks_to_copy=S.merge_dicts(to_copy, to_copy),
        ignored_seq_groups=[],
        num_lookahead_slots=self._get_num_lookahead_slots(is_prefill=False),
        need_score=self.need_score,
        allow_both_swap=True,
    )
    # ---- idle / runs bookkeeping for every live request ----------------------------------------
    # (scheduler.py:1358-1365: a request that ran has idle = 0 and, while promoted, one run less;
    # every other live request has been idle one tick more)
    for q in (swapped, running, waiting):
        for r in q:
            r.idle += 1
    for sg in out.scheduled_seq_groups:
        r = sg.seq_group
        if r.pri == -1:
            r.runs -= 1
        r.idle = 0
    return out


# ---------------------------------------------------------------------------------------------
# schedule(): SequenceGroupMetadata of the scheduled requests      (scheduler.py:1622-1678)
# ---------------------------------------------------------------------------------------------
def schedule(self):
    """``Scheduler.schedule`` for the configuration the LTR engine runs (no prefix caching, real
    allocation): the same ``(seq_group_metadata_list, scheduler_outputs)``, object for object.
    The reference calls, per scheduled request and per step, four block-manager methods that are
    no-ops without prefix caching (``access_all_blocks_in_seq`` :1652, ``get_common_computed_
    block_ids`` :1654 -> ``[]``, ``mark_blocks_as_computed`` :1690) and walks ``get_seqs`` twice;
    here the loop reads what it needs directly.  With prefix caching or ``fake_allocate`` the
    reference method runs unchanged."""
    bm = self.block_manager
    if getattr(bm, "enable_caching", False) or self.fake_allocate or \
            not hasattr(bm, "block_tables"):
        return self._vltr_ref_schedule()
    S = _ref()
    RUNNING = S.SequenceStatus.RUNNING
    SGM = S.SequenceGroupMetadata
    outs = self._schedule()
    now = time.time()
    tables = bm.block_tables
    need_score = outs.need_score
    any_prefill = outs.num_prefill_groups > 0
    metadata = []
    for ssg in outs.scheduled_seq_groups:
        sg = ssg.seq_group
        m = sg.metrics
        if m.first_scheduled_time is None:          # maybe_set_first_scheduled_time (:1634)
            m.first_scheduled_time = now
            m.time_in_queue = now - m.arrival_time
        seq_data = {}
        block_tables = {}
        for seq in sg.seqs_dict.values():
            if seq.status == RUNNING:
                sid = seq.seq_id
                seq_data[sid] = seq.data
                block_tables[sid] = [b.block_number for b in tables[sid]]
        metadata.append(SGM(
            request_id=sg.request_id, is_prompt=sg.is_prefill(), seq_data=seq_data,
            sampling_params=sg.sampling_params, block_tables=block_tables,
            token_chunk_size=ssg.token_chunk_size, lora_request=sg.lora_request,
            computed_block_nums=[], state=sg.state,
            multi_modal_data=sg.multi_modal_data if any_prefill else None,
            need_score=need_score))
    return metadata, outs


def install(scheduler, ranker=None):
    """Bind the B200 LTR tick on a reference ``Scheduler`` whose schedule_type is "opt*".
    ``ranker`` is for tests of the host logic only (an object with ``order(scores, pri, negate)``);
    the product always uses the CUDA ranker and raises without a GPU."""
    st = scheduler.schedule_type
    if not st.startswith("opt"):
        raise ValueError(f"vllm_ltr_b200 LTR tick handles schedule_type 'opt*', got {st!r}")
    scheduler._vltr_ranker = GpuRanker() if ranker is None else ranker
    scheduler._schedule = types.MethodType(general_schedule, scheduler)
    scheduler._get_ordered_requests = types.MethodType(get_opt_ordered_requests, scheduler)
    scheduler._update_priority = types.MethodType(update_opt_priority, scheduler)
    if not hasattr(scheduler, "_vltr_ref_schedule"):
        scheduler._vltr_ref_schedule = scheduler.schedule      # the reference's bound method
    scheduler.schedule = types.MethodType(schedule, scheduler)
    return scheduler
