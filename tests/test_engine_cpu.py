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
