"""CPU: the scheduler plugin inside the REAL reference ``LLMEngine`` (fake executor, see
tests/engine_fake.py).  The same seeded trace is run twice through ``LLMEngine.step()`` - once
with the reference's own LTR methods, once after ``install_engine()`` - and every step's
scheduled (request, chunk) list, every request's tokens and the finishing order must agree."""
import pytest

from tests import engine_fake as F
from tests import sched_replay as R

pytestmark = pytest.mark.skipif(not R.have_reference(),
                                reason="needs the reference's Python package")


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    R.import_reference()
    E = F.engine_trace()
    E._transformers5_compat()
    root = tmp_path_factory.mktemp("assets")
    assets = E.make_assets(str(root), "llama-tiny", "125m", layers=1, with_predictor=False)
    return E, assets


def _run(E, assets, plugin, monkeypatch, schedule, n=300, stop=0, blocks=600):
    llm = F.build_fake_llm(assets, schedule, gpu_blocks=blocks, cpu_blocks=32768)
    reqs = E.make_trace(n, 32000, 4096, max_output=48)
    eng = llm.llm_engine
    eng.scheduler.aux_model = R.StubScorer({})        # replaced by TimedScores' synthetic scores
    eng.scheduler.aux_model.obtain_aux_scores = lambda sgs: [
        sg.set_aux_model_score(0.0) for sg in sgs]
    if plugin:
        from vllm_ltr_b200 import integration, ltr_scheduler
        monkeypatch.setattr(ltr_scheduler, "GpuRanker", lambda: R.NumpyRanker())
        done = integration.install_engine(eng, predictor=False, prefill=False)
        assert done["scheduler"].startswith("LTR tick"), done
    out = E.run_trace(llm, reqs, scores="synthetic", record_schedule=True)
    return out


@pytest.mark.parametrize("schedule,stop", [("opt-xxx", 0), ("opt-xxx-synthetic((-1,-1,60),)", 60),
                                           ("opt-starv10-period2", 0)])
def test_plugin_matches_reference_scheduler_through_engine(env, monkeypatch, schedule, stop):
    E, assets = env
    a = _run(E, assets, False, monkeypatch, schedule)
    b = _run(E, assets, True, monkeypatch, schedule)
    sa, sb = a["summary"], b["summary"]
    assert sa["steps"] == sb["steps"] and sa["finished"] == sb["finished"]
    assert sa["generated_tokens"] == sb["generated_tokens"] > 0
    if stop:
        assert sa["finished"] < 300          # the synthetic stop rule cut the run short
    else:
        assert sa["finished"] == 300
    assert a["detail"]["schedule"] == b["detail"]["schedule"]
    assert a["detail"]["tokens"] == b["detail"]["tokens"]
    # pressure on the 600-block pool really forced preemption by swap in this scenario
    assert any(len(step) < 256 for step in a["detail"]["schedule"])


def test_install_engine_reads_the_reference_predictor_config(env, monkeypatch):
    """prefill_predictor_model_config is a PrefillPredictorConfig with one attribute,
    `.model` (config_predictor.py:131-143); install_engine must read .model.path/.pred_model/
    .max_length like llm_engine.py:229-236."""
    import types
    E, assets = env
    from vllm.config_predictor import PrefillPredictorConfig

    from vllm_ltr_b200 import integration, predictor
    cfg = PrefillPredictorConfig.from_dict({"model": {
        "pred_model": "/tok", "num_labels": 1, "mtype": "opt", "activation": "none",
        "path": "/ckpt", "max_length": 1536, "max_batch_size": 128}})
    seen = {}

    def fake(path, tokenizer=None, max_length=2048, **kw):
        seen.update(path=path, tokenizer=tokenizer, max_length=max_length, **kw)
        return "B200AUX"

    monkeypatch.setattr(predictor.B200AuxModel, "from_pretrained", staticmethod(fake))
    sched = types.SimpleNamespace(schedule_type="fcfs", aux_model=object())
    eng = types.SimpleNamespace(scheduler=sched, model_executor=None,
                                model_config=types.SimpleNamespace(
                                    prefill_predictor_model_config=cfg))
    done = integration.install_engine(eng, scheduler=False, prefill=False)
    assert sched.aux_model == "B200AUX" and done["predictor"].startswith("B200AuxModel")
    assert seen == {"path": "/ckpt", "tokenizer": "/tok", "max_length": 1536}


def test_trace_with_poisson_arrivals_through_engine(env, monkeypatch):
    """tools/engine_trace.run_trace with arrival times (the "Poisson" configuration,
    benchmark_serving_real.py:159-176): requests enter the engine as their arrival time passes,
    every one of them finishes with all its tokens, and latency counts from the arrival."""
    import numpy as np
    E, assets = env
    n = 120
    reqs = E.make_trace(n, 32000, 4096, max_output=24)
    arrivals = E.make_arrivals(n, rate=400.0)
    assert arrivals[0] == 0.0 and np.all(np.diff(arrivals) >= 0) and 0.1 < arrivals[-1] < 1.0
    assert abs(np.diff(E.make_arrivals(20000, rate=50.0)).mean() - 1 / 50.0) < 1e-3
    res = {}
    for plugin in (False, True):
        llm = F.build_fake_llm(assets, "opt-xxx", gpu_blocks=600, cpu_blocks=32768)
        eng = llm.llm_engine
        eng.scheduler.aux_model = R.StubScorer({})
        eng.scheduler.aux_model.obtain_aux_scores = lambda sgs: [
            sg.set_aux_model_score(0.0) for sg in sgs]
        if plugin:
            from vllm_ltr_b200 import integration, ltr_scheduler
            monkeypatch.setattr(ltr_scheduler, "GpuRanker", lambda: R.NumpyRanker())
            integration.install_engine(eng, predictor=False, prefill=False)
        seen_live = []
        inner = eng.scheduler.schedule

        def spy(inner=inner, eng=eng):
            seen_live.append(eng.scheduler.get_num_unfinished_seq_groups())
            return inner()
        eng.scheduler.schedule = spy
        out = E.run_trace(llm, reqs, scores="synthetic", arrivals=arrivals)["summary"]
        res[plugin] = out
        assert out["finished"] == n and out["generated_tokens"] == sum(r[2] for r in reqs)
        assert out["arrivals"].startswith(f"{n} requests over")
        assert out["elapsed_s"] >= arrivals[-1]
        # latency is measured from each request's own arrival: nothing can take longer than the run
        assert 0 < out["p50_latency_s"] <= out["p90_latency_s"] <= out["elapsed_s"]
        # the requests trickled in: the first steps saw only a few of them
        assert seen_live[0] < n // 2 and max(seen_live) > seen_live[0]
    assert res[False]["generated_tokens"] == res[True]["generated_tokens"]


def test_engine_trace_cli_with_request_rate(env, monkeypatch, tmp_path):
    """The command line bench.py's `trace_poisson` leg issues (tools/engine_trace.py --request-rate
    --request-time --max-output --stop-after 0), with the engine construction replaced by the fake
    executor: the arrival window cuts the trace, every kept request finishes, the JSON carries
    what bench.py reads."""
    import json
    import sys
    E, assets = env
    made = {}

    def fake_build(arm, a, num_stop, **kw):
        made["stop"] = num_stop
        llm = F.build_fake_llm(assets, "opt-xxx", gpu_blocks=600, cpu_blocks=32768)
        llm.llm_engine.scheduler.aux_model = R.StubScorer({})
        llm.llm_engine.scheduler.aux_model.obtain_aux_scores = lambda sgs: [
            sg.set_aux_model_score(0.0) for sg in sgs]
        return llm, {"arm": arm, "init_s": 0.0, "num_gpu_blocks": 600, "num_cpu_blocks": 32768}

    monkeypatch.setattr(E, "bind_arm", lambda arm: None)
    monkeypatch.setattr(E, "make_assets", lambda *a, **k: assets)
    monkeypatch.setattr(E, "build_llm", fake_build)
    monkeypatch.setitem(E.LLAMA, "llama-2-7b", dict(E.LLAMA["llama-tiny"]))
    out = tmp_path / "poisson.json"
    monkeypatch.setattr(sys, "argv", [
        "engine_trace.py", "--arm", "b200", "--model", "llama-2-7b", "--json", str(out),
        "--num-prompts", "400", "--stop-after", "0", "--request-rate", "300",
        "--request-time", "0.5", "--max-output", "16"])
    E.main()
    d = json.loads(out.read_text())
    s = d["summary"]
    assert made["stop"] == 0
    assert 60 < s["num_requests"] < 300 and s["finished"] == s["num_requests"]
    assert s["arrivals"].startswith(f"{s['num_requests']} requests over 0.")
    for k in ("tokens_per_s", "p50_latency_s", "p90_latency_s", "mean_latency_s", "requests_per_s",
              "generated_tokens", "elapsed_s", "steps", "schedule_s", "execute_model_s",
              "predictor_s", "other_host_s", "add_requests_s"):
        assert k in s, k
    assert d["engine"]["num_gpu_blocks"] == 600 and d["config"]["request_rate"] == 300.0
