"""One call that applies every seam of INTEGRATION.md to a built reference engine.

    import vllm_ltr_b200
    vllm_ltr_b200.install()                 # before `import vllm`: vllm._C -> libvllm_ltr_b200.so
    from vllm import LLM
    llm = LLM(...)                          # the unmodified reference engine
    from vllm_ltr_b200.integration import install_engine
    print(install_engine(llm.llm_engine))   # scheduler tick, predictor, swap path, decode inputs,
                                            # sampler, lm_head logits, attention backend,
                                            # prefill kernel

Each piece is optional and reported; pieces whose preconditions do not hold (schedule_type is not
"opt*", no predictor configured, Ray workers instead of an in-process driver worker) are skipped
with the reason, never silently replaced by something else.
"""
from typing import Dict, Optional


def _release_aux_engine(aux) -> None:
    """Drop the reference's second engine (AUXLLM -> AUXLLMEngine -> its own worker and a copy of
    the OPT weights on the GPU, llm_engine.py:228-242) once B200AuxModel has taken its place."""
    eng = getattr(aux, "llm_engine", None)
    worker = getattr(getattr(eng, "model_executor", None), "driver_worker", None)
    runner = getattr(worker, "model_runner", None)
    if runner is not None and hasattr(runner, "model"):
        runner.model = None
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    except ImportError:
        pass


def install_engine(engine, scheduler: bool = True, predictor: bool = True, swap: bool = True,
                   decode_inputs: bool = True, sampler: bool = True, prefill: bool = True,
                   attention_backend: bool = True, logits: bool = True,
                   step_broadcast: bool = True,
                   sampler_seed: Optional[int] = None,
                   release_reference_predictor: bool = True) -> Dict[str, str]:
    done: Dict[str, str] = {}
    sched = getattr(engine, "scheduler", None)

    if scheduler:
        st = getattr(sched, "schedule_type", "") if sched is not None else ""
        if sched is not None and st.startswith("opt"):
            from . import ltr_scheduler
            ltr_scheduler.install(sched)
            done["scheduler"] = ("LTR tick: CUDA rank kernel + host logic per running request "
                                 "(scheduler.py:969-1502)")
        else:
            done["scheduler"] = f"skipped: schedule_type {st!r} is not an LTR ('opt*') policy"

    if predictor:
        # model_config.prefill_predictor_model_config is a PrefillPredictorConfig whose only
        # attribute is `.model` (a PrefillModelConfig, vllm/config_predictor.py:38-46,131-143);
        # the reference reads .model.path / .model.pred_model / .model.max_length
        # (llm_engine.py:225-242).
        cfg = getattr(getattr(engine, "model_config", None), "prefill_predictor_model_config", None)
        m = getattr(cfg, "model", None)
        if sched is not None and m is not None and getattr(sched, "aux_model", None) is not None:
            from .predictor import B200AuxModel
            old = sched.aux_model
            sched.aux_model = B200AuxModel.from_pretrained(
                m.path, tokenizer=m.pred_model, max_length=m.max_length)
            if release_reference_predictor:
                _release_aux_engine(old)
            done["predictor"] = "B200AuxModel (tcgen05 OPT forward) behind scheduler.aux_model"
        else:
            done["predictor"] = "skipped: no prefill predictor configured on this engine"

    worker = getattr(getattr(engine, "model_executor", None), "driver_worker", None)
    runner = getattr(worker, "model_runner", None)
    if worker is None or runner is None:
        why = "skipped: no in-process driver worker (apply on every worker process instead)"
        for k, on in (("swap", swap), ("decode_inputs", decode_inputs), ("sampler", sampler)):
            if on:
                done[k] = why
    else:
        if swap:
            ce = getattr(worker, "cache_engine", None)
            if ce is not None:
                from ._C.cache_ops import install_cache_engine
                install_cache_engine(ce)
                done["swap"] = "CacheEngine.swap_in/out: all layers in one gather/scatter launch"
            else:
                done["swap"] = "skipped: cache engine not initialised yet"
        if decode_inputs:
            from . import decode_inputs as mr
            mr.install(runner)
            done["decode_inputs"] = ("ModelRunner._prepare_decode / _prepare_prompt: vectorised "
                                     "builders, one metadata copy per step")
        if sampler:
            model = getattr(runner, "model", None)
            if model is not None and hasattr(model, "sampler"):
                from .sampler import B200Sampler
                old = model.sampler
                if isinstance(old, B200Sampler):
                    old = old.reference
                model.sampler = B200Sampler(seed=sampler_seed, reference=old)
                done["sampler"] = ("B200Sampler (greedy / temperature requests; a step with any "
                                   "other request goes to the engine's own Sampler)")
            else:
                done["sampler"] = "skipped: model has no .sampler attribute"

    if logits:
        model = getattr(runner, "model", None)
        if model is None:
            done["logits"] = "skipped: no in-process model"
        else:
            from . import logits as lg
            if lg.install(model):
                done["logits"] = ("LogitsProcessor._get_logits: lm_head matmul on the tcgen05 "
                                  "dense layer (vltr_lm_head_logits)")
            else:
                done["logits"] = "skipped: model has no f16 / bf16 lm_head + logits_processor"

    if attention_backend:
        model = getattr(runner, "model", None)
        if model is not None:
            try:
                from . import attention_backend as ab
                n = ab.install(model)
                done["attention_backend"] = (f"B200AttentionImpl on {n} Attention layers: cache "
                                             "append, prefill and decode attention issued directly")
            except ImportError as e:
                done["attention_backend"] = f"skipped: {e}"
        else:
            done["attention_backend"] = "skipped: no in-process model"

    if step_broadcast:
        try:
            from . import step_broadcast as sb
            mods = sb.install()
            done["step_broadcast"] = ("broadcast_tensor_dict: one coalesced buffer per call "
                                      f"(bound in {len(mods)} modules; a bypass at world size 1)")
        except ImportError as e:
            done["step_broadcast"] = f"skipped: {e}"

    if prefill:
        try:
            from . import prefix_prefill
            prefix_prefill.install(engine_path=True)
            done["prefill"] = "context_attention_fwd (Triton) -> vltr_context_attention_fwd"
        except ImportError as e:      # the reference's attention package is not importable here
            done["prefill"] = f"skipped: {e}"
    return done
