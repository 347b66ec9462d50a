"""GPU: the COMPOSED LTR tick the product ships - vllm_ltr_b200.ltr_scheduler installed on the
reference ``Scheduler`` WITH the CUDA rank kernel (``GpuRanker`` -> ``vltr_ltr_rank_host``), not a
Python stand-in - replayed tick by tick against the reference class's own methods over every
scenario of tests/test_ltr_scheduler.py.  The reference package is the unmodified copy shipped to
the GPU box (oracle/_ref/vllm_py, ``make -C oracle vllm_py``)."""
import pytest

from tests import sched_replay as R
from tests.test_ltr_scheduler import SCENARIOS

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.have_reference(), reason="reference package not shipped")]


@pytest.fixture(scope="module")
def ranker():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from vllm_ltr_b200.ltr_scheduler import GpuRanker
    return GpuRanker()


@pytest.mark.parametrize("n,seed,stype,gpu,cpu,mt,ms", SCENARIOS)
def test_tick_by_tick_parity_with_cuda_ranker(ranker, n, seed, stype, gpu, cpu, mt, ms):
    ref = R.run(n, seed, stype, gpu, cpu, max_tokens=mt, max_seqs=ms, use_plugin=False)
    ours = R.run(n, seed, stype, gpu, cpu, max_tokens=mt, max_seqs=ms, use_plugin=True,
                 ranker=ranker)
    assert len(ref) > 5 and len(ours) == len(ref)
    for t, (a, b) in enumerate(zip(ref, ours)):
        for key in a:
            assert a[key] == b[key], f"tick {t}: {key} differs"


def test_synthetic_bound_with_cuda_ranker(ranker):
    st = "opt-xxx-synthetic((-1,-1,10),)"
    ref = R.run(60, 7, st, 4096, 4096, use_plugin=False)
    ours = R.run(60, 7, st, 4096, 4096, use_plugin=True, ranker=ranker)
    assert ref == ours and len(ref) > 3


def test_ten_thousand_live_requests_with_cuda_ranker(ranker):
    """synthetic-sharegpt.sh scale: 10 000 requests queued at t=0, the reference's 768-token /
    256-sequence budget, fp16-quantised scores (ties): 12 ticks, identical decisions."""
    import numpy as np
    R.import_reference()
    from vllm import SamplingParams
    from vllm.config import CacheConfig, SchedulerConfig
    from vllm.core.scheduler import Scheduler
    from vllm.sequence import Logprob, Sequence, SequenceGroup, SequenceStatus

    from vllm_ltr_b200 import ltr_scheduler

    def build(plugin):
        sc = SchedulerConfig(768, 256, 8192, enable_chunked_prefill=True, schedule_type="opt-xxx")
        cc = CacheConfig(16, 1.0, 1, "auto")
        cc.num_gpu_blocks, cc.num_cpu_blocks = 3000, 20000
        s = Scheduler(sc, cc, None)
        rng = np.random.default_rng(0)
        score = rng.normal(0, 1, 10000).astype(np.float16).astype(float)
        s.aux_model = R.StubScorer({str(i): float(score[i]) for i in range(10000)})
        if plugin:
            ltr_scheduler.install(s, ranker=ranker)
        for i in range(10000):
            toks = list(range(int(rng.integers(4, 200))))
            s.add_seq_group(SequenceGroup(str(i), [Sequence(i, "", toks, 16)],
                                          SamplingParams(ignore_eos=True, max_tokens=64), 0.0))
        return s

    outs = {}
    for plugin in (False, True):
        s = build(plugin)
        log = []
        for _ in range(12):
            _, o = s.schedule()
            log.append(([(x.seq_group.request_id, x.token_chunk_size)
                         for x in o.scheduled_seq_groups], len(o.blocks_to_swap_out),
                        len(o.blocks_to_swap_in), [g.request_id for g in s.running]))
            for x in o.scheduled_seq_groups:
                x.seq_group.update_num_computed_tokens(x.token_chunk_size)
                if x.seq_group.get_num_uncomputed_tokens() == 0:
                    for seq in x.seq_group.get_seqs(status=SequenceStatus.RUNNING):
                        seq.append_token_id(7, {7: Logprob(0.0)})
        outs[plugin] = log
    assert outs[False] == outs[True]
