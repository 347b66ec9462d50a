"""Host logic of the LTR scheduling tick (CPU): vllm_ltr_b200.ltr_scheduler installed on the
reference's Scheduler must reproduce, tick by tick, what the reference's own
_general_schedule / reserve_free_blocks / _get_opt_ordered_requests do
(vllm/core/scheduler.py:969-1000, 1101-1502): scheduled groups and chunk sizes, block tables,
swap-in/out/copy maps, queue orders and the idle/runs/pri side effects.

The reference has no test of this path ("LTR-specific: none"), so the pin
is the reference class itself, imported from /root/reference (present in the build container
only - these tests skip elsewhere).  The CUDA ranker is replaced by a Python `sorted` double here;
its own bit-exactness is tests/test_gpu_ltr.py."""
import pytest

from tests import sched_replay as R

pytestmark = pytest.mark.skipif(not R.have_reference(),
                                reason="needs the reference checkout at /root/reference")

SCENARIOS = [
    # (n, seed, schedule_type, gpu blocks, cpu blocks, max_tokens, max_seqs)
    (40, 0, "opt-xxx", 4096, 4096, 64, 16),            # no memory pressure
    (120, 1, "opt-xxx", 160, 4096, 64, 16),            # forced swap-outs and proactive swap-ins
    (120, 2, "opt-xxx-starv8-period4", 160, 4096, 64, 16),   # starvation control (pri)
    (200, 3, "opt-xxx", 96, 4096, 48, 8),              # heavy pressure: pinned requests dropped
    (64, 4, "opt-xxx-starv3-period2", 4096, 4096, 32, 4),
    (300, 5, "opt-xxx", 400, 4096, 768, 256),          # the reference's token / seq budgets
]


@pytest.mark.parametrize("n,seed,stype,gpu,cpu,mt,ms", SCENARIOS)
def test_tick_by_tick_parity(n, seed, stype, gpu, cpu, mt, ms):
    ref = R.run(n, seed, stype, gpu, cpu, max_tokens=mt, max_seqs=ms, use_plugin=False)
    ours = R.run(n, seed, stype, gpu, cpu, max_tokens=mt, max_seqs=ms, use_plugin=True,
                 ranker=R.PythonRanker())
    assert len(ref) > 5
    assert len(ours) == len(ref)
    for t, (a, b) in enumerate(zip(ref, ours)):
        for key in a:
            assert a[key] == b[key], f"tick {t}: {key} differs"
    # the scenario exercised what it was built for
    if gpu < 1000:
        assert any(r["swap_out"] for r in ref), "scenario never swapped out"
        assert any(r["swap_in"] for r in ref), "scenario never swapped in"
    if "starv" in stype:
        assert any(s[3] == -1 for r in ref for s in r["state"]), "no request was ever promoted"


def _same(ref, ours):
    assert len(ours) == len(ref) and len(ref) > 5
    for t, (a, b) in enumerate(zip(ref, ours)):
        for key in a:
            assert a[key] == b[key], f"tick {t}: {key} differs"


@pytest.mark.parametrize("stype", ["opt-xxx", "opt-xxx-starv6-period3"])
def test_parity_with_aborts(stype):
    """Requests aborted between ticks - from the middle of the waiting queue, from the running and
    the swapped queue, and ids that are not queued at all: the scores the plugin keeps for the
    waiting queue must be dropped and rebuilt, tick for tick like the reference."""
    aborts = {3: [30, 31, 70], 4: [119], 9: [0, 1, 2, 3, 4, 5], 10: [55], 11: [56, 57, 999],
              20: list(range(60, 100)), 26: [10, 11, 12]}
    kw = dict(max_tokens=64, max_seqs=16, aborts=aborts)
    ref = R.run(120, 11, stype, 160, 4096, use_plugin=False, **kw)
    ours = R.run(120, 11, stype, 160, 4096, use_plugin=True, ranker=R.PythonRanker(), **kw)
    _same(ref, ours)
    assert any(r["swap_out"] for r in ref)


def test_parity_when_many_waiting_requests_are_admitted_at_once(monkeypatch):
    """More waiting requests admitted in one tick than are deleted from the queue one by one: the
    queue (and the scores kept for it) is rebuilt instead."""
    from vllm_ltr_b200 import ltr_scheduler
    monkeypatch.setattr(ltr_scheduler, "_EDIT_IN_PLACE_MAX", 3)
    kw = dict(max_tokens=4096, max_seqs=256, max_prompt=12)
    ref = R.run(400, 13, "opt-xxx", 4096, 4096, use_plugin=False, **kw)
    ours = R.run(400, 13, "opt-xxx", 4096, 4096, use_plugin=True, ranker=R.NumpyRanker(), **kw)
    _same(ref, ours)
    assert max(r["n_prefill"] for r in ref) > 3


def test_parity_with_a_foreign_order():
    """Somebody else's ``_get_ordered_requests`` bound on the scheduler (here: the reference's own
    method): the tick takes whatever list it returns."""
    import types

    def run(plugin):
        hooked = []

        def on_tick(sched, groups, tick):
            if plugin and not hooked:
                from vllm.core.scheduler import Scheduler
                sched._get_ordered_requests = types.MethodType(
                    Scheduler._get_opt_ordered_requests, sched)
                hooked.append(1)
        return R.run(120, 3, "opt-xxx", 160, 4096, max_tokens=64, max_seqs=16, use_plugin=plugin,
                     ranker=R.PythonRanker(), on_tick=on_tick)
    _same(run(False), run(True))


def test_idle_counter_reads_like_the_eager_one():
    """Without starvation control the plugin counts ``idle`` by a per-scheduler tick clock while a
    request is queued; reads and writes behave like the plain attribute, and a request that left
    the scheduler (finished or aborted) is a plain ``SequenceGroup`` again."""
    R.import_reference()
    from vllm.sequence import SequenceGroup
    seen = {"queued": 0, "left": 0}

    def on_tick(sched, groups, tick):
        live = {id(g) for q in (sched.waiting, sched.running, sched.swapped) for g in q}
        for g in groups:
            if id(g) in live:
                assert type(g) is not SequenceGroup and isinstance(g, SequenceGroup)
                assert "idle" not in g.__dict__
                before = g.idle
                g.idle += 0                      # read-modify-write through the property
                assert g.idle == before
                seen["queued"] += 1
            elif hasattr(g, "pri"):              # was added, has left
                assert type(g) is SequenceGroup
                assert isinstance(g.__dict__["idle"], int)
                assert not any(k.startswith("_vltr") for k in g.__dict__)
                seen["left"] += 1
    aborts = {5: [7, 8], 12: [20]}
    ref = R.run(80, 21, "opt-xxx", 160, 4096, max_tokens=64, max_seqs=16, aborts=aborts)
    ours = R.run(80, 21, "opt-xxx", 160, 4096, max_tokens=64, max_seqs=16, aborts=aborts,
                 use_plugin=True, ranker=R.PythonRanker(), on_tick=on_tick)
    _same(ref, ours)
    assert seen["queued"] > 100 and seen["left"] > 100


def test_synthetic_bound_termination():
    """schedule_type "...synthetic((lo,hi,count),)" stops the run after `count` completions
    (scheduler.py:276-282,1697-1721) - handled by the reference's free_finished_seq_groups, which
    the plugin leaves in place."""
    st = "opt-xxx-synthetic((-1,-1,10),)"
    ref = R.run(60, 7, st, 4096, 4096, use_plugin=False)
    ours = R.run(60, 7, st, 4096, 4096, use_plugin=True, ranker=R.PythonRanker())
    assert ref == ours and len(ref) > 3


def test_install_rejects_other_policies():
    R.import_reference()
    from vllm.config import CacheConfig, SchedulerConfig
    from vllm.core.scheduler import Scheduler
    from vllm_ltr_b200 import ltr_scheduler
    sc = SchedulerConfig(64, 16, 4096, enable_chunked_prefill=True, schedule_type="fcfs")
    cc = CacheConfig(4, 1.0, 1, "auto")
    cc.num_gpu_blocks = cc.num_cpu_blocks = 64
    with pytest.raises(ValueError):
        ltr_scheduler.install(Scheduler(sc, cc, None), ranker=R.PythonRanker())


def test_product_ranker_requires_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vllm_ltr_b200 import ltr_scheduler
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ltr_scheduler.GpuRanker()


def test_tick_cost_is_linear():
    """10 000 queued requests: the reference tick is ~140 ms of Python;
    the plugin's host logic must be several times cheaper with identical output."""
    import time
    R.import_reference()
    import numpy as np
    from vllm import SamplingParams
    from vllm.config import CacheConfig, SchedulerConfig
    from vllm.core.scheduler import Scheduler
    from vllm.sequence import Logprob, Sequence, SequenceGroup, SequenceStatus
    from vllm_ltr_b200 import ltr_scheduler

    def build(plugin):
        sc = SchedulerConfig(768, 256, 8192, enable_chunked_prefill=True, schedule_type="opt-xxx")
        cc = CacheConfig(16, 1.0, 1, "auto")
        cc.num_gpu_blocks, cc.num_cpu_blocks = 20000, 20000
        s = Scheduler(sc, cc, None)
        rng = np.random.default_rng(0)
        score = rng.normal(0, 1, 10000).astype(np.float16).astype(float)
        s.aux_model = R.StubScorer({str(i): float(score[i]) for i in range(10000)})
        if plugin:
            ltr_scheduler.install(s, ranker=R.PythonRanker())
        for i in range(10000):
            toks = list(range(int(rng.integers(4, 200))))
            g = SequenceGroup(str(i), [Sequence(i, "", toks, 16)],
                              SamplingParams(ignore_eos=True, max_tokens=64), 0.0)
            s.add_seq_group(g)
        return s

    res = {}
    for plugin in (False, True):
        s = build(plugin)
        ticks, outs = [], []
        for _ in range(24):     # reach the steady state: a full running queue
            t0 = time.perf_counter()
            metas, o = s.schedule()
            ticks.append(time.perf_counter() - t0)
            outs.append([(x.seq_group.request_id, x.token_chunk_size) for x in o.scheduled_seq_groups])
            for x in o.scheduled_seq_groups:
                x.seq_group.update_num_computed_tokens(x.token_chunk_size)
                if x.seq_group.get_num_uncomputed_tokens() == 0:
                    for seq in x.seq_group.get_seqs(status=SequenceStatus.RUNNING):
                        seq.append_token_id(7, {7: Logprob(0.0)})
        res[plugin] = (min(ticks[-5:]), outs)
    assert res[False][1] == res[True][1]
    print(f"tick @10k live: reference {res[False][0]*1e3:.1f} ms, plugin {res[True][0]*1e3:.1f} ms")
    assert res[True][0] < res[False][0] / 2


def test_numpy_ranker_equals_python_sorted():
    """The stand-in ranker used by the CPU-side tools orders exactly like the reference's `sorted`
    expression (scheduler.py:996/998), ties included."""
    import numpy as np
    rng = np.random.default_rng(0)
    scores = rng.normal(0, 1, 5000).astype(np.float16).astype(np.float64).tolist()   # many ties
    pri = rng.choice([-1, 0], 5000).tolist()
    a, b = R.NumpyRanker(), R.PythonRanker()
    assert a.order(scores).tolist() == b.order(scores).tolist()
    assert a.order(scores, pri).tolist() == b.order(scores, pri).tolist()
    assert a.order(scores, None, negate=False).tolist() == b.order(scores, None, negate=False).tolist()
