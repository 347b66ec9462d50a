"""CPU: install_engine() wires every seam onto a reference-shaped engine object and reports what
it did; preconditions that do not hold are reported as skipped (never silently substituted)."""
import types

import pytest

from tests import sched_replay as R

pytestmark = pytest.mark.skipif(not R.have_reference(),
                                reason="needs the reference checkout at /root/reference")


def test_install_engine_on_reference_shaped_objects(monkeypatch):
    R.import_reference()
    from vllm.config import CacheConfig, SchedulerConfig
    from vllm.core.scheduler import Scheduler

    from vllm_ltr_b200 import decode_inputs, integration, ltr_scheduler
    sc = SchedulerConfig(768, 256, 8192, enable_chunked_prefill=True, schedule_type="opt-xxx")
    cc = CacheConfig(16, 1.0, 1, "auto")
    cc.num_gpu_blocks, cc.num_cpu_blocks = 64, 64
    sched = Scheduler(sc, cc, None)
    # no GPU here: the ranker constructor is the only piece that needs one
    monkeypatch.setattr(ltr_scheduler, "GpuRanker", lambda: R.PythonRanker())
    runner = types.SimpleNamespace(model=types.SimpleNamespace(sampler=object()))
    ce = types.SimpleNamespace(cpu_cache=[], gpu_cache=[], swap_in=None, swap_out=None)
    engine = types.SimpleNamespace(
        scheduler=sched, model_config=types.SimpleNamespace(prefill_predictor_model_config=None),
        model_executor=types.SimpleNamespace(
            driver_worker=types.SimpleNamespace(model_runner=runner, cache_engine=ce)))
    done = integration.install_engine(engine, prefill=False)
    assert done["scheduler"].startswith("LTR tick")
    assert sched._schedule.__func__ is ltr_scheduler.general_schedule
    assert done["predictor"].startswith("skipped")
    assert callable(ce.swap_in) and callable(ce.swap_out)
    assert runner._prepare_decode.__func__ is decode_inputs.prepare_decode
    assert type(runner.model.sampler).__name__ == "B200Sampler"

    fcfs = types.SimpleNamespace(scheduler=types.SimpleNamespace(schedule_type="fcfs"),
                                 model_config=None, model_executor=None)
    done = integration.install_engine(fcfs, prefill=False)
    assert done["scheduler"].startswith("skipped") and done["swap"].startswith("skipped")
    assert done["sampler"].startswith("skipped") and done["decode_inputs"].startswith("skipped")
