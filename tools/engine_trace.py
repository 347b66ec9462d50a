#!/usr/bin/env python
"""Run the UNMODIFIED reference engine (vllm 0.4.1 fork, shipped as oracle/_ref/vllm_py by
`make -C oracle vllm_py`) over the synthetic ShareGPT-shaped trace of BASELINE.md section 4, in
one of two arms, and report the reference's own headline metric (output tokens/s and p50 request
latency, benchmarks/benchmark_throughput_original.py:251-270,436-442).

    arm "ref_cuda" : vllm._C := oracle/_ref/ref_cuda_C (the reference's CUDA kernels recompiled
                     for sm_100a), stock Python scheduler, stock AUXLLM predictor engine
    arm "b200"     : vllm_ltr_b200.install() before `import vllm`, then
                     vllm_ltr_b200.integration.install_engine(llm.llm_engine)

Everything outside those two seams is identical: same engine code, same flags
(--enable-chunked-prefill --enforce-eager, 768-token / 256-sequence budget, block 16, schedule
type "opt-xxx-synthetic((-1,-1,N),)"), same seeded weights, same prompts.

TEST / BENCH INFRASTRUCTURE.  There is no network, so nothing can be downloaded: model
directories (config.json + a word-level tokenizer whose vocabulary is "t0".."t<V-1>"), the
OPT-125M predictor checkpoint (seeded random weights written as safetensors in the layout
train/trainer.py saves) and the trace are generated locally.  Weights of the main model are the
engine's `--load-format dummy` tensors re-drawn from a seeded N(0, 0.02) (norm weights 1) so
that logits are not degenerate; both arms draw the same values.

    python tools/engine_trace.py --arm b200 --model llama-2-7b --num-prompts 10000 \
        --stop-after 1000 --json gpurun_out/trace_b200.json
"""
import argparse
import json
import math
import os
import sys
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VLLM_PY = os.path.join(ROOT, "oracle", "_ref", "vllm_py")

LLAMA = {
    # hidden, inter, layers, q heads, kv heads, vocab, max positions, rope theta, dtype
    "llama-2-7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=32, vocab_size=32000,
                       max_position_embeddings=4096, rope_theta=10000.0, rms_norm_eps=1e-5,
                       torch_dtype="float16"),
    "llama-3-8b": dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=8, vocab_size=128256,
                       max_position_embeddings=8192, rope_theta=500000.0, rms_norm_eps=1e-5,
                       torch_dtype="bfloat16"),
    # test-sized: same head geometry as Llama-2-7B, fewer / narrower layers
    "llama-tiny": dict(hidden_size=1024, intermediate_size=2816, num_hidden_layers=4,
                       num_attention_heads=8, num_key_value_heads=8, vocab_size=32000,
                       max_position_embeddings=4096, rope_theta=10000.0, rms_norm_eps=1e-5,
                       torch_dtype="float16"),
}
OPT = {
    "125m": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, ffn_dim=3072,
                 word_embed_proj_dim=768, do_layer_norm_before=True),
    "350m": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, ffn_dim=4096,
                 word_embed_proj_dim=512, do_layer_norm_before=False),
}
OPT_VOCAB = 50272


# --------------------------------------------------------------------------------------------
# assets: model dirs, tokenizers, predictor checkpoint
# --------------------------------------------------------------------------------------------
def _write_tokenizer(path: str, vocab: int, bos: int, eos: int, prefix_bos: bool) -> None:
    """A fast (tokenizers-library) word-level tokenizer: token i is the word "t<i>"; ids 0 / bos /
    eos are "<unk>" / "<s>" / "</s>" (special-token strings must not be prefixes of words)."""
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    words = {f"t{i}": i for i in range(vocab)}
    names = {0: "<unk>", bos: "<s>", eos: "</s>"}
    for i, n in names.items():
        del words[f"t{i}"]
        words[n] = i
    tok = Tokenizer(models.WordLevel(words, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    if prefix_bos:
        tok.post_processor = processors.TemplateProcessing(
            single=f"{names[bos]} $A", special_tokens=[(names[bos], bos)])
    tok.save(os.path.join(path, "tokenizer.json"))
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "bos_token": names[bos],
                   "eos_token": names[eos], "unk_token": "<unk>", "model_max_length": 1 << 30,
                   "clean_up_tokenization_spaces": False}, f)
    with open(os.path.join(path, "special_tokens_map.json"), "w") as f:
        json.dump({"bos_token": names[bos], "eos_token": names[eos], "unk_token": "<unk>"}, f)


def make_assets(root: str, model: str = "llama-2-7b", predictor: str = "125m",
                layers: Optional[int] = None, pred_max_length: int = 2048,
                seed: int = 0, with_predictor: bool = True) -> Dict[str, str]:
    """Create (idempotently) the main-model dir, the predictor dir and its usage_config.json."""
    import torch
    os.makedirs(root, exist_ok=True)
    cfg = dict(LLAMA[model])
    if layers is not None:
        cfg["num_hidden_layers"] = layers
    tag = f"{model}-L{cfg['num_hidden_layers']}"
    mdir = os.path.join(root, tag)
    if not os.path.exists(os.path.join(mdir, "tokenizer.json")):
        os.makedirs(mdir, exist_ok=True)
        with open(os.path.join(mdir, "config.json"), "w") as f:
            json.dump({"architectures": ["LlamaForCausalLM"], "model_type": "llama",
                       "hidden_act": "silu", "bos_token_id": 1, "eos_token_id": 2,
                       "tie_word_embeddings": False, "attention_bias": False,
                       "initializer_range": 0.02, "use_cache": True, **cfg}, f)
        _write_tokenizer(mdir, cfg["vocab_size"], 1, 2, prefix_bos=True)

    if not with_predictor:
        return {"model": mdir, "predictor": "", "usage_config": ""}
    pdir = os.path.join(root, f"opt-{predictor}-ltr-seed{seed}")
    if not os.path.exists(os.path.join(pdir, "model.safetensors")):
        os.makedirs(pdir, exist_ok=True)
        o = OPT[predictor]
        with open(os.path.join(pdir, "config.json"), "w") as f:
            json.dump({"architectures": ["OPTForSequenceClassification"], "model_type": "opt",
                       "activation_function": "relu", "vocab_size": OPT_VOCAB,
                       "max_position_embeddings": 2048, "enable_bias": True,
                       "layer_norm_elementwise_affine": True, "_remove_final_layer_norm": False,
                       "bos_token_id": 2, "eos_token_id": 2, "pad_token_id": 1, "num_labels": 1,
                       "id2label": {"0": "LABEL_0"}, "label2id": {"LABEL_0": 0},
                       "dropout": 0.0, "torch_dtype": "float16", **o}, f)
        _write_tokenizer(pdir, OPT_VOCAB, 2, 2, prefix_bos=True)
        from safetensors.torch import save_file
        save_file(opt_state_dict(predictor, seed), os.path.join(pdir, "model.safetensors"),
                  metadata={"format": "pt"})
    ucfg = os.path.join(pdir, "usage_config.json")
    with open(ucfg, "w") as f:      # PrefillPredictorConfig.from_json, config_predictor.py:131-143
        json.dump({"model": {"pred_model": pdir, "num_labels": 1, "mtype": "opt",
                             "activation": "none", "path": pdir, "max_length": pred_max_length,
                             "max_batch_size": 512}}, f)
    return {"model": mdir, "predictor": pdir, "usage_config": ucfg}


def opt_state_dict(kind: str, seed: int):
    """Seeded random OPT classifier weights under the names HF's OPTForSequenceClassification
    (the class train/trainer.py saves) uses; fp16, N(0, 0.02) matrices, LayerNorm ~ 1."""
    import torch
    o = OPT[kind]
    h, L, ffn, e = o["hidden_size"], o["num_hidden_layers"], o["ffn_dim"], o["word_embed_proj_dim"]
    g = torch.Generator().manual_seed(seed)

    def mat(*shape, std=0.02):
        return (torch.randn(*shape, generator=g) * std).half()

    sd = {"model.decoder.embed_tokens.weight": mat(OPT_VOCAB, e),
          "model.decoder.embed_positions.weight": mat(2048 + 2, h),
          "score.weight": mat(1, e, std=0.2)}
    if e != h:
        sd["model.decoder.project_in.weight"] = mat(h, e)
        sd["model.decoder.project_out.weight"] = mat(e, h)
    if o["do_layer_norm_before"]:
        sd["model.decoder.final_layer_norm.weight"] = (1 + mat(h, std=0.05))
        sd["model.decoder.final_layer_norm.bias"] = mat(h)
    for i in range(L):
        p = f"model.decoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = mat(h, h, std=0.04)
            sd[p + f"self_attn.{n}.bias"] = mat(h)
        sd[p + "self_attn_layer_norm.weight"] = (1 + mat(h, std=0.05))
        sd[p + "self_attn_layer_norm.bias"] = mat(h)
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = mat(ffn, h, std=0.04), mat(ffn)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = mat(h, ffn, std=0.04), mat(h)
        sd[p + "final_layer_norm.weight"] = (1 + mat(h, std=0.05))
        sd[p + "final_layer_norm.bias"] = mat(h)
    return {k: v.contiguous() for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
# the trace
# --------------------------------------------------------------------------------------------
def make_trace(n: int, vocab: int, max_len: int, seed: int = 0, tau: float = 0.52,
               max_output: Optional[int] = None, min_output: Optional[int] = None):
    """[(prompt text, prompt_len, output_len, ordering score)]; prompt_len counts the BOS the
    tokenizer prepends.  Scores are -log(output_len) + N(0, sigma), sigma bisected so that Kendall
    tau(score, -output_len) = `tau`, rounded to fp16 (forces ties)."""
    import scipy.stats
    rng = np.random.default_rng(seed)
    plen = np.clip(np.round(np.exp(rng.normal(5.0, 1.0, n))), 4, 2048).astype(np.int64)
    olen = np.round(np.exp(rng.normal(5.3, 1.1, n))).astype(np.int64)
    olen = np.clip(olen, 4, np.minimum(8192, max_len) - plen)
    if min_output is not None:
        olen = np.maximum(olen, min_output)
    if max_output is not None:
        olen = np.minimum(olen, max_output)
    noise = rng.normal(0.0, 1.0, n)
    base = -np.log(olen.astype(np.float64))
    lo, hi = 0.0, 20.0
    sub = slice(0, min(n, 4000))
    for _ in range(30):
        mid = 0.5 * (lo + hi)
        t = scipy.stats.kendalltau(base[sub] + mid * noise[sub], -olen[sub]).statistic
        lo, hi = (mid, hi) if t > tau else (lo, mid)
    score = (base + 0.5 * (lo + hi) * noise).astype(np.float16).astype(np.float64)
    reqs = []
    for i in range(n):
        ids = rng.integers(1000, vocab - 1000, int(plen[i]) - 1)
        reqs.append((" ".join(f"t{t}" for t in ids), int(plen[i]), int(olen[i]), float(score[i])))
    return reqs


def make_arrivals(n: int, rate: float, cv: float = 1.0, seed: int = 0) -> np.ndarray:
    """Arrival offsets in seconds of `n` requests sent at `rate` requests/s: the first at t=0, then
    Gamma(shape=1/cv^2, scale=cv^2/rate) gaps - benchmark_serving_real.py:159-176 (cv=1: Poisson
    arrivals, the "Poisson" configuration of bench-lmsys.sh)."""
    rng = np.random.default_rng(seed + 1)
    gaps = rng.gamma(1.0 / (cv * cv), cv * cv / rate, max(n - 1, 0))
    return np.concatenate(([0.0], np.cumsum(gaps)))[:n]


# --------------------------------------------------------------------------------------------
# engine construction
# --------------------------------------------------------------------------------------------
def _load_ref_cuda():
    import glob
    import importlib.util

    import torch  # noqa: F401
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_cuda_C.*.so"))
    assert hits, "oracle/_ref/ref_cuda_C not built (make -C oracle ref_cuda)"
    spec = importlib.util.spec_from_file_location("ref_cuda_C", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def bind_arm(arm: str) -> None:
    """Choose what `from vllm._C import ops` (vllm/_custom_ops.py:5-9) resolves to; must run
    before `import vllm`."""
    assert "vllm" not in sys.modules, "bind_arm() must run before `import vllm`"
    assert os.path.isdir(os.path.join(VLLM_PY, "vllm")), \
        "oracle/_ref/vllm_py missing (make -C oracle vllm_py, needs /root/reference)"
    sys.path.insert(0, VLLM_PY)
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
    os.environ.setdefault("VLLM_DO_NOT_TRACK", "1")
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
    if arm == "b200":
        import vllm_ltr_b200
        vllm_ltr_b200.install()
    elif arm == "ref_cuda":
        ref = _load_ref_cuda()
        sys.modules["vllm._C"] = ref
        for sub in ("ops", "cache_ops", "cuda_utils", "custom_ar"):
            sys.modules[f"vllm._C.{sub}"] = getattr(ref, sub)
    else:
        raise ValueError(arm)
    import vllm.utils
    # the box has no route to the outside: the engine's get_ip() probe (utils.py) cannot work
    vllm.utils.get_ip = lambda: "127.0.0.1"
    import vllm.executor.gpu_executor as ge
    ge.get_ip = vllm.utils.get_ip
    import vllm  # noqa: F401
    if arm == "ref_cuda":
        setattr(sys.modules["vllm"], "_C", sys.modules["vllm._C"])
    _transformers5_compat()


def _transformers5_compat() -> None:
    """The reference pins transformers 4.40; the image has 5.x, whose LlamaConfig always carries
    `rope_scaling` (= rope_parameters, {"rope_type": "default", "rope_theta": ...}) and no
    top-level `rope_theta`.  Present the 4.x view to the engine (both arms alike)."""
    import vllm.config as vc
    inner = vc.get_config

    def get_config(*a, **k):
        cfg = inner(*a, **k)
        rp = getattr(cfg, "rope_parameters", None) or getattr(cfg, "rope_scaling", None)
        if isinstance(rp, dict) and rp.get("rope_type", rp.get("type", "default")) == "default":
            theta = rp.get("rope_theta", None)
            try:
                cfg.rope_scaling = None
            except Exception:
                pass
            if getattr(cfg, "rope_scaling", None) is not None:
                object.__setattr__(cfg, "rope_scaling", None)
            if theta is not None and not isinstance(getattr(cfg, "rope_theta", None), (int, float)):
                cfg.rope_theta = float(theta)
        return cfg

    vc.get_config = get_config
    from transformers import PreTrainedTokenizerBase
    if not hasattr(PreTrainedTokenizerBase, "all_special_tokens_extended"):
        # removed in transformers 5; read by get_cached_tokenizer (transformers_utils/tokenizer.py:29)
        PreTrainedTokenizerBase.all_special_tokens_extended = property(
            lambda self: list(self.all_special_tokens))


def reinit_weights(model, seed: int = 0, std: float = 0.02) -> None:
    """Replace the dummy loader's U(-1e-3, 1e-3) tensors by seeded N(0, std) (norm weights 1).
    Drawn parameter by parameter in name order from one seeded generator on the parameters' own
    device, so both arms (same box, same GPU type) get the same values."""
    import torch
    params = sorted(model.named_parameters(), key=lambda kv: kv[0])
    dev = params[0][1].device
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.no_grad():
        for name, p in params:
            if name.endswith("norm.weight") or "layernorm.weight" in name:
                p.fill_(1.0)
            else:
                p.normal_(0.0, std, generator=g)


def build_llm(arm: str, assets: Dict[str, str], num_stop: int, schedule: str = "opt-xxx",
              kv_cache_dtype: str = "auto", gpu_memory_utilization: float = 0.9,
              swap_space: int = 16, max_model_len: Optional[int] = None, seed: int = 0,
              predictor: bool = True, install_opts: Optional[dict] = None,
              num_gpu_blocks_override: Optional[int] = None):
    """LLM(...) with the flags of benchmarks/synthetic-sharegpt.sh:4; returns (llm, info)."""
    from vllm import LLM
    st = schedule if num_stop <= 0 else f"{schedule}-synthetic((-1,-1,{num_stop}),)"
    t0 = time.perf_counter()
    kw = {}
    if num_gpu_blocks_override is not None:
        kw["num_gpu_blocks_override"] = num_gpu_blocks_override
    llm = LLM(model=assets["model"], tokenizer=assets["model"], seed=seed,
              load_format="dummy", enforce_eager=True, enable_chunked_prefill=True,
              max_num_batched_tokens=768, max_num_seqs=256, block_size=16,
              kv_cache_dtype=kv_cache_dtype, schedule_type=st, swap_space=swap_space,
              gpu_memory_utilization=gpu_memory_utilization, max_model_len=max_model_len,
              prefill_predictor_model_config=assets["usage_config"] if predictor else "",
              disable_log_stats=True, **kw)
    eng = llm.llm_engine
    info = {"arm": arm, "schedule_type": st, "init_s": None,
            "num_gpu_blocks": eng.cache_config.num_gpu_blocks,
            "num_cpu_blocks": eng.cache_config.num_cpu_blocks}
    reinit_weights(eng.model_executor.driver_worker.model_runner.model, seed)
    if arm == "b200":
        from vllm_ltr_b200.integration import install_engine
        info["installed"] = install_engine(eng, **(install_opts or {}))
    info["init_s"] = time.perf_counter() - t0
    return llm, info


class TimedScores:
    """Wraps scheduler.aux_model: the real predictor still runs on the prompt tokens (and is
    timed); with `synthetic` given, its scores are then replaced by the trace's ordering scores
     so that both arms schedule on identical values."""

    def __init__(self, inner, synthetic: Optional[Dict[str, float]]):
        self.inner, self.synthetic = inner, synthetic
        self.seconds, self.calls, self.prompts = 0.0, 0, 0
        self.predicted: Dict[str, float] = {}

    def obtain_aux_scores(self, seq_groups):
        t0 = time.perf_counter()
        ret = self.inner.obtain_aux_scores(seq_groups)
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        self.prompts += len(seq_groups)
        for sg in seq_groups:
            self.predicted[sg.request_id] = float(sg.aux_model_score)
        if self.synthetic is not None:
            ret = []
            for sg in seq_groups:
                sg.set_aux_model_score(self.synthetic[sg.request_id])
                ret.append(self.synthetic[sg.request_id])
        return ret

    def __getattr__(self, k):
        return getattr(self.inner, k)


def run_trace(llm, reqs, scores: str = "synthetic", record_schedule: bool = False,
              logprobs: Optional[int] = None, temperature: float = 0.0,
              max_steps: Optional[int] = None, profile: Optional[tuple] = None,
              arrivals: Optional[Sequence[float]] = None) -> dict:
    """The request loop of benchmark_throughput_original.py:251-270 + LLM._run_engine
    (entrypoints/llm.py:221-246), instrumented per step.  With `arrivals` (seconds from the start,
    ascending, one per request) a request enters the engine at the first step boundary after its
    arrival time - what the API server's engine loop does with requests that come in while a
    step runs (async_llm_engine.py:209) - and its latency counts from its arrival."""
    import torch
    from vllm import SamplingParams
    eng = llm.llm_engine
    sched = eng.scheduler
    synth = {}

    def params(output_len):
        return SamplingParams(n=1, temperature=temperature, top_p=1.0, use_beam_search=False,
                              ignore_eos=True, max_tokens=output_len, est_tokens=None,
                              logprobs=logprobs)

    t_add = 0.0
    arrived_at: Dict[str, float] = {}
    rids = [str(next(llm.request_counter)) for _ in reqs]
    for rid, (_, _, _, score) in zip(rids, reqs):
        synth[rid] = score
    if arrivals is None:
        t_add = time.perf_counter()
        for rid, (prompt, _, output_len, _) in zip(rids, reqs):
            eng.add_request(rid, prompt, params(output_len), None)
        t_add = time.perf_counter() - t_add
    else:
        assert len(arrivals) == len(reqs)
        assert all(b >= a for a, b in zip(arrivals, arrivals[1:])), "arrivals must ascend"
    aux = None
    if getattr(sched, "aux_model", None) is not None:
        aux = TimedScores(sched.aux_model, synth if scores == "synthetic" else None)
        sched.aux_model = aux

    schedule_log: List[list] = []
    t_sched = t_exec = 0.0
    inner_schedule = sched.schedule
    exe = eng.model_executor
    inner_execute = exe.execute_model

    def timed_schedule():
        nonlocal t_sched
        t = time.perf_counter()
        out = inner_schedule()
        t_sched += time.perf_counter() - t
        if record_schedule:
            schedule_log.append([(s.seq_group.request_id, s.token_chunk_size)
                                 for s in out[1].scheduled_seq_groups])
        return out

    def timed_execute(*a, **k):
        nonlocal t_exec
        t = time.perf_counter()
        out = inner_execute(*a, **k)
        t_exec += time.perf_counter() - t
        return out

    sched.schedule = timed_schedule
    exe.execute_model = timed_execute
    outputs, steps = [], 0
    sync = torch.cuda.synchronize if torch.cuda.is_available() else (lambda: None)
    sync()
    start = time.perf_counter()
    sched.start_time = time.time()
    prof = None
    nxt = 0
    n_req = len(reqs)
    while True:
        if arrivals is not None and nxt < n_req:
            now = time.perf_counter() - start
            if arrivals[nxt] <= now:
                t = time.perf_counter()
                while nxt < n_req and arrivals[nxt] <= now:
                    prompt, _, output_len, _ = reqs[nxt]
                    eng.add_request(rids[nxt], prompt, params(output_len), None)
                    arrived_at[rids[nxt]] = float(arrivals[nxt])
                    nxt += 1
                t_add += time.perf_counter() - t
            elif not eng.has_unfinished_requests():
                time.sleep(min(max(arrivals[nxt] - now, 0.0), 0.05))     # idle until the next arrival
                continue
        if not eng.has_unfinished_requests():
            if arrivals is None or nxt >= n_req:
                break
            continue
        if profile is not None and steps == profile[0]:
            from torch.profiler import ProfilerActivity
            from torch.profiler import profile as _profile
            sync()
            prof = _profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA])
            prof.__enter__()
            prof_t0 = time.perf_counter()
        for out in eng.step():
            if out.finished:
                out.latency = time.perf_counter() - start - arrived_at.get(out.request_id, 0.0)
                outputs.append(out)
        steps += 1
        if prof is not None and steps == profile[0] + profile[1]:
            sync()
            wall = time.perf_counter() - prof_t0
            prof.__exit__(None, None, None)
            _dump_profile(prof, wall, profile[1], profile[2])
            prof = None
        if max_steps is not None and steps >= max_steps:
            break
    sync()
    elapsed = time.perf_counter() - start
    sched.schedule, exe.execute_model = inner_schedule, inner_execute
    if aux is not None:
        sched.aux_model = aux.inner
    outputs.sort(key=lambda o: int(o.request_id))
    gen = sum(len(o.outputs[0].token_ids) for o in outputs)
    lat = [o.latency for o in outputs]
    res = {
        "num_requests": len(reqs), "finished": len(outputs), "generated_tokens": gen,
        "arrivals": "all at t=0" if arrivals is None else
                    f"{len(reqs)} requests over {float(arrivals[-1]) if len(reqs) else 0.0:.1f} s",
        "elapsed_s": elapsed, "steps": steps,
        "tokens_per_s": gen / elapsed if elapsed > 0 else 0.0,
        "requests_per_s": len(outputs) / elapsed if elapsed > 0 else 0.0,
        "p50_latency_s": float(np.percentile(lat, 50)) if lat else None,
        "p90_latency_s": float(np.percentile(lat, 90)) if lat else None,
        "mean_latency_s": float(np.mean(lat)) if lat else None,
        "add_requests_s": t_add, "schedule_s": t_sched, "execute_model_s": t_exec,
        "other_host_s": elapsed - t_sched - t_exec,
        "predictor_s": aux.seconds if aux else 0.0,
        "predictor_calls": aux.calls if aux else 0,
        "predictor_prompts": aux.prompts if aux else 0,
        "scores": scores,
    }
    detail = {
        "tokens": {o.request_id: list(o.outputs[0].token_ids) for o in outputs},
        "schedule": schedule_log,
        "predicted_scores": aux.predicted if aux else {},
    }
    if logprobs:
        lp = {}
        for o in outputs:
            rows = []
            for d in (o.outputs[0].logprobs or []):
                rows.append({int(t): float(getattr(v, "logprob", v)) for t, v in d.items()})
            lp[o.request_id] = rows
        detail["logprobs"] = lp
    return {"summary": res, "detail": detail}


def _dump_profile(prof, wall_s, nsteps, path):
    """Per-kernel device time and the host-side picture of `nsteps` engine steps."""
    ka = prof.key_averages()
    dev = []
    for e in ka:
        t = getattr(e, "self_device_time_total", 0) or 0
        if t > 0:
            dev.append((t, e.count, e.key))
    dev.sort(reverse=True)
    busy = sum(t for t, _, _ in dev)
    with open(path, "w") as f:
        f.write(f"{nsteps} engine steps, wall {wall_s * 1e3:.1f} ms = {wall_s / nsteps * 1e3:.2f} "
                f"ms/step; device busy {busy / 1e3:.1f} ms ({100 * busy / 1e6 / wall_s:.0f}% of wall)\n\n")
        f.write("device time by kernel (self):\n")
        for t, c, k in dev[:45]:
            f.write(f"  {t / 1e3:9.2f} ms {100 * t / busy:5.1f}%  x{c:<6d} {k[:120]}\n")
        f.write("\nhost (CPU self time):\n")
        f.write(ka.table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=70))
    print(f"profile of {nsteps} steps -> {path}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["ref_cuda", "b200"], required=True)
    ap.add_argument("--model", default="llama-2-7b", choices=sorted(LLAMA))
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--predictor", default="125m", choices=sorted(OPT))
    ap.add_argument("--num-prompts", type=int, default=10000)
    ap.add_argument("--stop-after", type=int, default=1000,
                    help='N of "opt-xxx-synthetic((-1,-1,N),)"; 0 = run every request to the end')
    ap.add_argument("--max-output", type=int, default=None)
    ap.add_argument("--min-output", type=int, default=None)
    ap.add_argument("--request-rate", type=float, default=None,
                    help="requests/s with Gamma(cv) gaps (benchmark_serving_real.py:159-176) instead "
                         "of every request at t=0")
    ap.add_argument("--request-time", type=float, default=None,
                    help="with --request-rate: only the requests arriving within this many seconds")
    ap.add_argument("--cv", type=float, default=1.0)
    ap.add_argument("--scores", choices=["synthetic", "predictor"], default="synthetic")
    ap.add_argument("--kv-cache-dtype", default="auto")
    ap.add_argument("--gpu-memory-utilization", type=float, default=0.9)
    ap.add_argument("--swap-space", type=int, default=16)
    ap.add_argument("--num-gpu-blocks", type=int, default=None)
    ap.add_argument("--logprobs", type=int, default=None)
    ap.add_argument("--record", action="store_true", help="keep tokens / schedule in the JSON")
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--profile", default=None,
                    help="START:COUNT:FILE - torch.profiler over COUNT engine steps from step START")
    ap.add_argument("--assets", default="/tmp/vltr_assets")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()

    bind_arm(args.arm)
    assets = make_assets(args.assets, args.model, args.predictor, args.layers)
    cfg = dict(LLAMA[args.model])
    reqs = make_trace(args.num_prompts, cfg["vocab_size"], cfg["max_position_embeddings"],
                      max_output=args.max_output, min_output=args.min_output)
    llm, info = build_llm(args.arm, assets, args.stop_after, kv_cache_dtype=args.kv_cache_dtype,
                          gpu_memory_utilization=args.gpu_memory_utilization,
                          swap_space=args.swap_space,
                          num_gpu_blocks_override=args.num_gpu_blocks)
    arrivals = None
    if args.request_rate is not None:
        arrivals = make_arrivals(len(reqs), args.request_rate, args.cv)
        if args.request_time is not None:
            keep = int(np.searchsorted(arrivals, args.request_time, side="right"))
            reqs, arrivals = reqs[:keep], arrivals[:keep]
    out = run_trace(llm, reqs, scores=args.scores, record_schedule=args.record,
                    logprobs=args.logprobs, max_steps=args.max_steps, arrivals=arrivals,
                    profile=((int(args.profile.split(":")[0]), int(args.profile.split(":")[1]),
                              args.profile.split(":", 2)[2]) if args.profile else None))
    out["engine"] = info
    out["config"] = {k: v for k, v in vars(args).items()}
    if not args.record:
        out.pop("detail")
    line = json.dumps({"engine": info, **out["summary"]})
    print("TRACE " + line, flush=True)
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    main()
