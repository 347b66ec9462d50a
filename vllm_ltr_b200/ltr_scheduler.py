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
* the bookkeeping around it does not grow with the number of queued requests: the reference
  scans deques for membership (``x in deque``, :1139/:1166/:1187), removes the selected requests
  one ``list.remove`` at a time (:1213-1214), tests every live request against the list of
  scheduled ones (:1358-1365) and rebuilds three lists of all live requests per tick - about
  100-140 ms per tick at 10 000 live requests.  Here
  - the scores of the waiting queue (which never change once set) are carried from tick to tick
    as the float64 array the rank kernel reads: arrivals appended, admitted positions deleted
    (``_WaitingScores``; re-read from the requests whenever anything else touched the queue);
  - the ordered list of all live requests is never materialised: the tick walks the rank
    permutation, which queue a rank came from is a range test on its index, the admitted requests
    are a prefix of it and the eviction / swap-in candidates are masks over the rest;
  - the admitted waiting requests are deleted from the deque at their known positions;
  - without starvation control, ``idle += 1`` on every live request (:1365) is one increment of
    a tick clock that ``request.idle`` is read against (``_counting_class``).
  What remains per tick is the work on the <= max_num_seqs requests that actually run (the
  reference's own block-manager and budget calls) plus the rank kernel.

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
from itertools import chain, compress, islice
from operator import itemgetter
from typing import Dict, Iterable, List, Optional

import numpy as np

from ._lib import check, lib

_WAITING, _RUNNING, _SWAPPED = 0, 1, 2
_EDIT_IN_PLACE_MAX = 64      # admitted waiting requests deleted from the queue one by one up to here


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
# `SequenceGroup.idle` without a pass over every live request per tick
# ---------------------------------------------------------------------------------------------
class _TickClock:
    """Number of ticks this scheduler has run."""
    __slots__ = ("now",)

    def __init__(self):
        self.now = 0


_COUNTING: Dict[type, type] = {}


def _counting_class(base: type) -> type:
    """Subclass of the reference's ``SequenceGroup`` whose ``idle`` is computed on read.

    scheduler.py:1358-1365 ends every tick with ``idle += 1`` on every live request that did not
    run - at 10 000 queued requests that is 10 000 scattered attribute updates (cache misses on
    10 000 instance dicts) per engine step, for a counter nothing reads unless starvation
    control is on.  While a request is queued in a scheduler ticked by this plugin WITHOUT
    starvation control its class is this subclass: ``idle`` reads as
    ``value when last set + ticks since`` and the tick advances one clock instead.  Reads and
    writes of ``request.idle`` give exactly what the reference's eager counter would hold; when
    the request leaves the scheduler (finished, aborted, run cut off by the synthetic bound) it
    is turned back into a plain ``SequenceGroup`` with a plain ``idle`` attribute."""
    cls = _COUNTING.get(base)
    if cls is None:
        def _get(self):
            d = self.__dict__
            return d["_vltr_idle"] + d["_vltr_clock"].now - d["_vltr_since"]

        def _set(self, value):
            d = self.__dict__
            d["_vltr_idle"] = value
            d["_vltr_since"] = d["_vltr_clock"].now

        cls = type("TickCounted" + base.__name__, (base,),
                   {"__slots__": (), "idle": property(_get, _set), "__module__": __name__,
                    "__doc__": _counting_class.__doc__})
        globals()[cls.__name__] = cls              # importable by name (pickle)
        _COUNTING[base] = cls
    return cls


def _count_by_clock(req, clock: _TickClock) -> None:
    base = type(req)
    if base in _COUNTING.values():
        return
    d = req.__dict__
    d["_vltr_idle"] = d.pop("idle", 0)
    d["_vltr_since"] = clock.now
    d["_vltr_clock"] = clock
    req.__class__ = _counting_class(base)


def _count_eagerly(req) -> None:
    cls = type(req)
    if cls not in _COUNTING.values():
        return
    value = req.idle
    req.__class__ = cls.__mro__[1]
    d = req.__dict__
    del d["_vltr_idle"], d["_vltr_since"], d["_vltr_clock"]
    d["idle"] = value


def free_finished_seq_groups(self) -> None:
    """The reference's method (scheduler.py:1697-1723), then: requests that left the queues get
    their plain ``idle`` attribute back."""
    running, waiting, swapped = self.running, self.waiting, self.swapped
    n = len(running)
    self._vltr_ref_free_finished()
    if self.running is not running and len(self.running) != n:
        kept = set(map(id, self.running))
        for r in running:
            if id(r) not in kept:
                _count_eagerly(r)
    # the synthetic / time bound was reached: every queue replaced by an empty one
    if self.waiting is not waiting:
        for r in waiting:
            _count_eagerly(r)
    if self.swapped is not swapped:
        for r in swapped:
            _count_eagerly(r)


def abort_seq_group(self, request_id) -> None:
    """The reference's method (scheduler.py:379-413); the aborted requests get their plain
    ``idle`` attribute back first."""
    ids = {request_id} if isinstance(request_id, str) else set(request_id)
    for q in (self.waiting, self.running, self.swapped):
        for r in q:
            if r.request_id in ids:
                _count_eagerly(r)
    return self._vltr_ref_abort(ids)


# ---------------------------------------------------------------------------------------------
# _get_ordered_requests for schedule_type "opt*"        (scheduler.py:969-1000)
# ---------------------------------------------------------------------------------------------
class _WaitingScores:
    """Scores of the waiting queue kept between ticks (starvation control off).

    A waiting request's score never changes once the predictor has set it, and the reference
    only ever appends to the waiting queue (``add_seq_group``), removes from it
    (``abort_seq_group``) or - this tick itself - takes the admitted requests out.  So the float64
    array the rank kernel reads is carried over from tick to tick: new arrivals are appended, the
    admitted positions deleted, and no pass over the ~10^4 queued requests is needed to rebuild
    it.  ``valid_for`` is the check that nothing else touched the queue: same deque object, not
    shorter, and the first and last request of the remembered content still sit at index 0 and
    n-1 (any ``remove`` at or before n-1 shifts the last one to n-2)."""
    __slots__ = ("queue", "reqs", "scores")

    def __init__(self, queue, reqs, scores):
        self.queue, self.reqs, self.scores = queue, reqs, scores

    def valid_for(self, queue) -> bool:
        n = len(self.reqs)
        return (queue is self.queue and len(queue) >= n
                and (n == 0 or (queue[0] is self.reqs[0] and queue[n - 1] is self.reqs[-1])))


def _scores_of_waiting(self, reqs: list) -> list:
    """aux_model_score of every request of `reqs`, after sending the unscored ones to the
    predictor in queue order (scheduler.py:971-982).  need_aux_model_score() is
    `aux_model_score is None` (vllm/sequence.py:460-461): the attribute is read directly."""
    scores = [r.aux_model_score for r in reqs]
    if None in scores:
        unscored = [r for r, sc in zip(reqs, scores) if sc is None]
        timed = int(os.environ.get("OPT_TIME", 0))
        t0 = time.time() if timed else 0.0
        self.aux_model.obtain_aux_scores(unscored)
        if timed:
            print("OPT-TIME: ", time.time() - t0)
        scores = [r.aux_model_score for r in reqs]
    return scores


def _rank_live(self):
    """One ranking of the live requests.  Returns ``(wreqs, others, order, nw, nr)``:
    live = wreqs (the waiting queue, a list) ++ others (running ++ swapped), `order` the
    permutation of range(len(live)) in schedule order (int32), nw / nr the number of waiting /
    running requests.  Which queue a rank came from is a range test on `order`."""
    w = self.waiting
    nw = len(w)
    keep = self.starv == -1
    cache = getattr(self, "_vltr_wcache", None) if keep else None
    clock = self._vltr_clock
    if cache is not None and cache.valid_for(w):
        wreqs, wscores = cache.reqs, cache.scores
        n0 = len(wreqs)
        if nw > n0:                                      # arrivals since the last tick
            tail = [w[i] for i in range(n0 - nw, 0)]
            ts = np.asarray(_scores_of_waiting(self, tail), dtype=np.float64)
            wreqs.extend(tail)
            cache.scores = wscores = np.concatenate((wscores, ts))
            if clock is not None:
                for r in tail:
                    _count_by_clock(r, clock)
    else:
        wreqs = list(w)
        wscores = np.asarray(_scores_of_waiting(self, wreqs), dtype=np.float64)
        self._vltr_wcache = _WaitingScores(w, wreqs, wscores) if keep else None
        if clock is not None:
            for r in wreqs:
                _count_by_clock(r, clock)
    others = list(self.running)
    nr = len(others)
    others.extend(self.swapped)
    pri = None
    if not keep:
        # starvation control: a request idle for `starv` ticks is promoted for `period`
        # scheduled steps, then demoted again
        for r in chain(wreqs, others):
            if r.idle >= self.starv:
                r.pri, r.idle, r.runs = -1, 0, self.period
            elif r.pri == -1 and r.runs <= 0:
                r.pri = 0
        pri = [r.pri for r in chain(wreqs, others)]
    if others:
        scores = np.concatenate(
            (wscores, np.asarray([r.aux_model_score for r in others], dtype=np.float64)))
    else:
        scores = wscores
    order = self._vltr_ranker.order(scores, pri, negate=True)
    return wreqs, others, order, nw, nr


def get_opt_ordered_requests(self) -> list:
    """The reference's return value: every live request, best rank first.  (The tick below does
    not materialise this list - it walks the permutation.)"""
    wreqs, others, order, _, _ = _rank_live(self)
    live = wreqs + others
    if len(live) >= 2:
        return list(itemgetter(*order.tolist())(live))
    return live


def update_opt_priority(self) -> None:     # scheduler.py:1002-1003
    return None


def _foreign_order(self, ordered: list):
    """An order produced by somebody else's ``_get_ordered_requests``: expressed as the same
    ``(wreqs, others, order, nw, nr)`` by looking every request up by identity."""
    wreqs = list(self.waiting)
    others = list(self.running)
    nw, nr = len(wreqs), len(others)
    others.extend(self.swapped)
    pos = {id(r): i for i, r in enumerate(chain(wreqs, others))}
    try:
        order = np.fromiter((pos[id(r)] for r in ordered), dtype=np.int32, count=len(ordered))
    except KeyError:
        raise AssertionError("seqgroup not in all lists") from None    # scheduler.py:1211
    self._vltr_wcache = None
    if self._vltr_clock is not None:
        for r in chain(wreqs, others):
            _count_by_clock(r, self._vltr_clock)
    return wreqs, others, order, nw, nr


def _iter_ranks(order: np.ndarray, block: int = 512):
    for s in range(0, len(order), block):
        yield from order[s:s + block].tolist()


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
    try:
        return _tick(self)
    except BaseException:
        self._vltr_wcache = None          # whatever state the queues are left in is re-read
        raise


def _tick(self):
    S = _ref()
    WAITING, RUNNING, SWAPPED = (S.SequenceStatus.WAITING, S.SequenceStatus.RUNNING,
                                 S.SequenceStatus.SWAPPED)
    self._update_priority()
    if getattr(self._get_ordered_requests, "__func__", None) is get_opt_ordered_requests:
        wreqs, others, order, nw, nr = _rank_live(self)
    else:
        wreqs, others, order, nw, nr = _foreign_order(self, self._get_ordered_requests())
    nwr = nw + nr
    live_before = nw + len(others)
    cfg = self.scheduler_config
    budget = S.SchedulingBudget(token_budget=cfg.max_num_batched_tokens,
                                max_num_seqs=cfg.max_num_seqs)
    final_budget = S.SchedulingBudget(token_budget=cfg.max_num_batched_tokens,
                                      max_num_seqs=cfg.max_num_seqs)
    # which queue a request sits in: a range test on its index into wreqs ++ running ++ swapped
    # while the ranks are walked; afterwards running and swapped requests (at most max_num_seqs
    # plus whatever was preempted - hundreds) are looked up by identity and everything else is a
    # waiting request.  No pass over the ~10^4 waiting requests is needed anywhere below.
    where: Dict[int, int] = {id(r): _RUNNING for r in islice(others, nr)}
    for r in islice(others, nr, None):
        where[id(r)] = _SWAPPED

    # ---- admit in rank order until a budget runs out: the selection is a prefix of the order
    bm = self.block_manager
    blocks_needed = 0
    selected, selected_at = [], []
    for idx in _iter_ranks(order):
        if idx < nw:
            req = wreqs[idx]
            new_tokens = self._get_num_new_tokens(req, WAITING, True, budget)
            assert new_tokens <= self.prompt_limit, "req exceed prompt limit"
            new_seqs = req.get_max_num_running_seqs()
            if new_tokens == 0 or not budget.can_schedule(num_new_tokens=new_tokens,
                                                          num_new_seqs=new_seqs):
                break
            blocks = len(req.get_seqs()[0].logical_token_blocks)
        elif idx < nwr:
            req = others[idx - nw]
            new_tokens = self._get_num_new_tokens(req, RUNNING, True, budget)
            if new_tokens == 0:
                assert budget.remaining_token_budget() == 0
                break
            new_seqs = req.get_max_num_running_seqs()
            if not budget.can_schedule(num_new_tokens=new_tokens, num_new_seqs=new_seqs):
                break
            blocks = new_seqs
        else:
            req = others[idx - nw]
            new_seqs = req.get_max_num_running_seqs()
            new_tokens = self._get_num_new_tokens(req, SWAPPED, True, budget)
            n_swapped = req.num_seqs(status=SWAPPED)
            if new_tokens == 0 or not budget.can_schedule(num_new_tokens=new_tokens,
                                                          num_new_seqs=new_seqs):
                break
            blocks = len(bm._get_physical_blocks(req)) + n_swapped
        budget.add_num_batched_tokens(req.request_id, new_tokens)
        budget.add_num_seqs(req.request_id, new_seqs)
        req.num_new_tokens, req.num_new_seqs = new_tokens, new_seqs
        blocks_needed += blocks
        selected.append(req)
        selected_at.append(idx)
    # the unselected members of the running queue (lowest rank first) and of the swapped queue
    # (best rank first): range tests on the rest of the permutation
    tail = order[len(selected):]
    evict_at = tail[(tail >= nw) & (tail < nwr)][::-1].tolist()
    swapin_at = tail[tail >= nwr].tolist()
    evict_candidates = (others[i - nw] for i in evict_at)
    swapin_candidates = (others[i - nw] for i in swapin_at)

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
        tag = where.get(id(req), _WAITING)
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

    # ---- the three queues after the tick -----------------------------------------------------------
    gone_w, gone_r, gone_s = taken[_WAITING], taken[_RUNNING] | dropped_running, taken[_SWAPPED]
    waiting = self.waiting
    if gone_w:
        # the admitted waiting requests sit at known positions of the queue: their indices
        at = sorted((i for i, r in zip(selected_at, selected) if i < nw and id(r) in gone_w),
                    reverse=True)
        cache = getattr(self, "_vltr_wcache", None)
        if len(at) <= _EDIT_IN_PLACE_MAX:
            for i in at:                       # a handful per tick: deleted in place
                del waiting[i]
                del wreqs[i]
            if cache is not None:
                cache.scores = np.delete(cache.scores, at)
        else:
            keep = bytearray(b"\x01") * nw
            for i in at:
                keep[i] = 0
            wreqs = list(compress(wreqs, keep))
            waiting = deque(wreqs)
            if cache is not None:
                self._vltr_wcache = _WaitingScores(
                    waiting, wreqs, cache.scores[np.frombuffer(keep, dtype=np.bool_)])
    if preempted:
        waiting.extendleft(preempted)
        self._vltr_wcache = None
    running = deque(r for r in islice(others, nr) if id(r) not in gone_r)
    for group in (new_prefill, run_decode, run_prefill, swp_decode, swp_prefill):
        running.extend(s.seq_group for s in group)
    swapped = deque(r for r in islice(others, nr, None) if id(r) not in gone_s)
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
    clock = self._vltr_clock
    if clock is not None:
        clock.now += 1                    # every live request: one tick more (read on demand)
    else:
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
    scheduler._vltr_wcache = None
    # starvation control reads every request's idle count every tick: counted eagerly, as the
    # reference does; otherwise by the tick clock (see _counting_class)
    scheduler._vltr_clock = _TickClock() if scheduler.starv == -1 else None
    if scheduler._vltr_clock is not None:
        for r in chain(scheduler.waiting, scheduler.running, scheduler.swapped):
            _count_by_clock(r, scheduler._vltr_clock)
        if not hasattr(scheduler, "_vltr_ref_free_finished"):
            scheduler._vltr_ref_free_finished = scheduler.free_finished_seq_groups
            scheduler._vltr_ref_abort = scheduler.abort_seq_group
        scheduler.free_finished_seq_groups = types.MethodType(free_finished_seq_groups, scheduler)
        scheduler.abort_seq_group = types.MethodType(abort_seq_group, scheduler)
    scheduler._schedule = types.MethodType(general_schedule, scheduler)
    scheduler._get_ordered_requests = types.MethodType(get_opt_ordered_requests, scheduler)
    scheduler._update_priority = types.MethodType(update_opt_priority, scheduler)
    if not hasattr(scheduler, "_vltr_ref_schedule"):
        scheduler._vltr_ref_schedule = scheduler.schedule      # the reference's bound method
    scheduler.schedule = types.MethodType(schedule, scheduler)
    return scheduler
