"""The learning-to-rank predictor behind ``scheduler.aux_model``.

The reference scores every new prompt once with an OPT-125M/350M sequence classifier by running a
complete second engine (``AUXLLMEngine.obtain_aux_scores``, vllm/engine/aux_llm_engine.py:332-412:
re-tokenise, build SequenceGroups, schedule, execute ``OPTForSequenceClassification`` through a
worker, sample one dummy token, collect ``logits[:,0].tolist()``).  Only three things in that path
are observable from the scheduler: which token ids the predictor sees (the prompt re-encoded with
the predictor's tokenizer, truncated to ``max_model_len``), the model's logits, and the call
``seq_group.set_aux_model_score(float)``.  ``B200AuxModel`` reproduces exactly those, with the
forward pass done by ``libvllm_ltr_b200`` (csrc/opt_predictor.cu: tcgen05 GEMMs with fused
epilogues, one stream, no host sync until the scores are read).

    m = model_config.prefill_predictor_model_config.model          # PrefillModelConfig
    aux = B200AuxModel.from_pretrained(m.path, tokenizer=m.pred_model, max_length=m.max_length)
    engine.scheduler.aux_model = aux                                # llm_engine.py:228-242
"""
import ctypes
import json
import os
from itertools import chain
from typing import Callable, Dict, List, Optional, Sequence

import torch

from ._lib import check, declare, lib

MAX_LAYERS = 48
F16, BF16 = 1, 2


class _Config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("hidden", "layers", "heads", "ffn", "vocab", "max_pos", "embed_dim", "num_labels",
                 "do_layer_norm_before", "final_layer_norm", "dtype")]


_PER_LAYER = ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_w", "ln2_b", "fc1_w",
              "fc1_b", "fc2_w", "fc2_b")


class _Weights(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in
                 ("embed_tokens", "embed_positions", "project_in", "project_out", "final_ln_w",
                  "final_ln_b", "score")] +
                [(n, ctypes.c_void_p * MAX_LAYERS) for n in _PER_LAYER])


_P = ctypes.c_void_p
declare("vltr_opt_create", ctypes.c_int, [ctypes.POINTER(_Config), ctypes.POINTER(_Weights),
                                          ctypes.POINTER(ctypes.c_void_p)])
declare("vltr_opt_destroy", None, [_P])
declare("vltr_opt_workspace_bytes", ctypes.c_size_t, [_P, ctypes.c_int, ctypes.c_int])
declare("vltr_opt_forward", ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, _P, ctypes.c_size_t, _P, _P])
declare("vltr_opt_gemm", ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, _P])


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """act(a @ w.T + bias) + resid on the tcgen05 path (the predictor's dense layer)."""
    if not a.is_cuda:
        raise RuntimeError("vllm_ltr_b200 predictor needs CUDA tensors (no CPU path)")
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty(m, n, dtype=a.dtype, device=a.device)
    code = {torch.float16: F16, torch.bfloat16: BF16}.get(a.dtype)
    if code is None:
        raise RuntimeError(f"Unsupported data type: {a.dtype}")
    with torch.cuda.device(a.device):
        check(lib.vltr_opt_gemm(a.data_ptr(), w.data_ptr(), out.data_ptr(),
                                bias.data_ptr() if bias is not None else None,
                                resid.data_ptr() if resid is not None else None, m, n, k,
                                1 if relu else 0, code,
                                torch._C._cuda_getCurrentRawStream(a.device.index)))
    return out


def dummy_checkpoint(kind: str = "125m", seed: int = 0, num_labels: int = 1,
                     vocab: int = 50272, max_pos: int = 2048):
    """(config, state_dict) with seeded random weights in the OPT-125M / OPT-350M predictor layouts
    - the counterpart of the reference's ``--load-format dummy`` (model_loader/loader.py:246) for
    benchmarks without checkpoints."""
    if kind == "125m":
        cfg = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, ffn_dim=3072,
                   word_embed_proj_dim=768, do_layer_norm_before=True)
    elif kind == "350m":
        cfg = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, ffn_dim=4096,
                   word_embed_proj_dim=512, do_layer_norm_before=False)
    else:
        raise ValueError(kind)
    cfg.update(vocab_size=vocab, max_position_embeddings=max_pos, num_labels=num_labels,
               activation_function="relu", _remove_final_layer_norm=False)
    g = torch.Generator().manual_seed(seed)
    H, E, Fd = cfg["hidden_size"], cfg["word_embed_proj_dim"], cfg["ffn_dim"]
    w = lambda *shape, std=0.02: (torch.randn(*shape, generator=g) * std).to(torch.float16)
    sd = {"model.decoder.embed_tokens.weight": w(vocab, E),
          "model.decoder.embed_positions.weight": w(max_pos + 2, H),
          "score.weight": w(num_labels, E, std=0.05)}
    if E != H:
        sd["model.decoder.project_in.weight"] = w(H, E)
        sd["model.decoder.project_out.weight"] = w(E, H)
    if cfg["do_layer_norm_before"]:
        sd["model.decoder.final_layer_norm.weight"] = 1 + w(H, std=0.1)
        sd["model.decoder.final_layer_norm.bias"] = w(H)
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.decoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"], sd[p + f"self_attn.{n}.bias"] = w(H, H, std=0.04), w(H)
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = w(Fd, H, std=0.04), w(Fd)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = w(H, Fd, std=0.04), w(H)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = 1 + w(H, std=0.1), w(H)
    return cfg, sd


class OPTPredictor:
    """OPTForSequenceClassification forward (vllm/model_executor/models/opt.py:362-397) on the
    GPU.  ``state_dict`` uses the HF checkpoint names the reference's load_weights consumes
    (opt.py:411-444): ``model.decoder.layers.N.self_attn.q_proj.weight`` ... ``score.weight``."""

    def __init__(self, config: Dict, state_dict: Dict[str, torch.Tensor], device="cuda:0",
                 dtype: torch.dtype = torch.float16, max_batch_tokens: int = 16384,
                 max_batch_prompts: int = 2048):
        if not torch.cuda.is_available():
            raise RuntimeError("vllm_ltr_b200 predictor needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.dtype = dtype
        g = lambda k, d=None: config.get(k, d)
        self.hidden = g("hidden_size")
        self.layers = g("num_hidden_layers")
        self.heads = g("num_attention_heads")
        self.ffn = g("ffn_dim")
        self.vocab = g("vocab_size")
        self.max_pos = g("max_position_embeddings")
        self.embed_dim = g("word_embed_proj_dim", self.hidden)
        self.num_labels = g("num_labels", 1)
        self.pre_ln = bool(g("do_layer_norm_before", True))
        self.final_ln = self.pre_ln and not g("_remove_final_layer_norm", False)
        if g("activation_function", "relu") != "relu":
            raise RuntimeError("the LTR predictor checkpoints use ReLU (OPT); got "
                               f"{g('activation_function')}")
        # HF names: "model.decoder.*" + "score.weight"; older checkpoints drop the "model." prefix
        # (the reference's loader accepts both, opt.py:421-424)
        sd = {(k[len("model."):] if k.startswith("model.") else k): v
              for k, v in state_dict.items()}
        self._keep: List[torch.Tensor] = []

        def dev(name, optional=False):
            if name not in sd:
                if optional:
                    return None
                raise KeyError(f"predictor checkpoint lacks {name}")
            t = sd[name].detach().to(self.device, dtype).contiguous()
            self._keep.append(t)
            return t

        def ptr(t):
            return t.data_ptr() if t is not None else None

        w = _Weights()
        self.embed_tokens = dev("decoder.embed_tokens.weight")
        w.embed_tokens = ptr(self.embed_tokens)
        w.embed_positions = ptr(dev("decoder.embed_positions.weight"))
        w.project_in = ptr(dev("decoder.project_in.weight", True))
        w.project_out = ptr(dev("decoder.project_out.weight", True))
        if (self.embed_dim != self.hidden) != (w.project_in is not None):
            raise RuntimeError("project_in/project_out must exist iff word_embed_proj_dim != hidden")
        if self.final_ln:
            w.final_ln_w = ptr(dev("decoder.final_layer_norm.weight"))
            w.final_ln_b = ptr(dev("decoder.final_layer_norm.bias"))
        w.score = ptr(dev("score.weight"))
        for i in range(self.layers):
            p = f"decoder.layers.{i}."
            qkv_w = torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], dim=0)
            qkv_b = torch.cat([sd[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], dim=0)
            sd[p + "qkv.weight"], sd[p + "qkv.bias"] = qkv_w, qkv_b
            for field, name in (("ln1_w", "self_attn_layer_norm.weight"),
                                ("ln1_b", "self_attn_layer_norm.bias"),
                                ("qkv_w", "qkv.weight"), ("qkv_b", "qkv.bias"),
                                ("out_w", "self_attn.out_proj.weight"),
                                ("out_b", "self_attn.out_proj.bias"),
                                ("ln2_w", "final_layer_norm.weight"),
                                ("ln2_b", "final_layer_norm.bias"),
                                ("fc1_w", "fc1.weight"), ("fc1_b", "fc1.bias"),
                                ("fc2_w", "fc2.weight"), ("fc2_b", "fc2.bias")):
                getattr(w, field)[i] = ptr(dev(p + name))
        cfg = _Config(self.hidden, self.layers, self.heads, self.ffn, self.vocab, self.max_pos,
                      self.embed_dim, self.num_labels, int(self.pre_ln), int(self.final_ln),
                      F16 if dtype == torch.float16 else BF16)
        self._handle = ctypes.c_void_p()
        check(lib.vltr_opt_create(ctypes.byref(cfg), ctypes.byref(w), ctypes.byref(self._handle)))
        self._weights = w
        self.max_batch_tokens = max_batch_tokens
        self.max_batch_prompts = max_batch_prompts
        self._ws = None
        self._ws_cap = (0, 0)
        # pinned staging ring: the host packs batch k+1 (and k+2) while the copy of batch k waits
        # on the stream behind the forward of batch k-1
        self._stage = [[None, None] for _ in range(3)]        # [buffer, copy-done event]
        self._stage_i = 0
        # single-prompt forwards (the steady-state tick scores one new prompt) replay a CUDA graph
        # per 64-token bucket: ~90 launches and 2 tensor-map encodes per dense layer become one
        # graph launch.  Keyed by padded token count; dropped when the workspace is reallocated.
        self.use_graphs = os.environ.get("VLTR_OPT_GRAPHS", "1") != "0"
        self._graphs = {}

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            lib.vltr_opt_destroy(h)
            self._handle = None

    @classmethod
    def from_pretrained(cls, path: str, **kw) -> "OPTPredictor":
        """Load an HF OPTForSequenceClassification checkpoint directory (what train/trainer.py
        saves, :214-226): config.json + model.safetensors | pytorch_model.bin."""
        with open(os.path.join(path, "config.json")) as f:
            config = json.load(f)
        st = os.path.join(path, "model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu",
                            weights_only=True)
        return cls(config, sd, **kw)

    def _workspace(self, tokens: int, prompts: int):
        if self._ws is None or tokens > self._ws_cap[0] or prompts > self._ws_cap[1]:
            cap = (max(tokens, self._ws_cap[0], 256), max(prompts, self._ws_cap[1], 16))
            n = lib.vltr_opt_workspace_bytes(self._handle, cap[0], cap[1])
            self._ws = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
            self._ws_cap = cap
            self._graphs.clear()          # captured launches point into the old workspace
        off = (-self._ws.data_ptr()) % 256
        return self._ws.data_ptr() + off, self._ws.numel() - off

    def logits(self, prompts: Sequence[Sequence[int]]) -> torch.Tensor:
        """[len(prompts), num_labels] float32 logits (values already rounded to the model dtype),
        on the device.  Prompts are packed into varlen batches of <= max_batch_tokens tokens and
        <= max_batch_prompts prompts per forward.  (The reference's aux scheduler batches
        <= max(max_model_len, 2048) tokens and <= 256 prompts per step, config.py:586; a prompt's
        score does not depend on what it is batched with, and larger batches fill the dense-layer
        grids: 128-row tiles x 148 SMs.)"""
        out = torch.empty(len(prompts), self.num_labels, dtype=torch.float32, device=self.device)
        i = 0
        while i < len(prompts):
            j, tok = i, 0
            while (j < len(prompts) and j - i < self.max_batch_prompts
                   and (j == i or tok + len(prompts[j]) <= self.max_batch_tokens)):
                tok += len(prompts[j])
                j += 1
            self._forward(prompts[i:j], out[i:j])
            i = j
        return out

    def _forward_graph(self, prompt: Sequence[int], out: torch.Tensor) -> None:
        """One prompt through the captured forward of its 64-token bucket.  Rows past the prompt
        are padding (token 0 at position 0): every op of the forward is row-wise except attention,
        which reads cu_seqlens = [0, len] and never touches them."""
        L = len(prompt)
        mp = (L + 63) // 64 * 64
        need = 2 * mp + 3                                  # ids | positions | cu_seqlens | last
        ent = self._graphs.get(mp)
        if ent is None:
            ws, ws_bytes = self._workspace(max(mp, 2048), 16)
            ent = {"dev": torch.zeros(need, dtype=torch.int32, device=self.device),
                   "host": torch.zeros(need, dtype=torch.int32).pin_memory(),
                   "out": torch.empty(1, self.num_labels, dtype=torch.float32, device=self.device),
                   "evt": None, "graph": None}
            base = ent["dev"].data_ptr()

            def call():
                check(lib.vltr_opt_forward(self._handle, base, base + 4 * mp, base + 8 * mp,
                                           base + 8 * mp + 8, mp, 1, mp, ws,
                                           ctypes.c_size_t(ws_bytes), ent["out"].data_ptr(),
                                           torch._C._cuda_getCurrentRawStream(self.device.index)))
            ent["dev"][2 * mp + 1] = 1                     # a valid 1-token prompt for the warm-up
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.device(self.device), torch.cuda.stream(side):
                call()
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.device(self.device), torch.cuda.graph(g):
                call()
            ent["graph"] = g
            self._graphs[mp] = ent
        if ent["evt"] is not None:
            ent["evt"].synchronize()                       # previous copy out of the pinned buffer
        h = ent["host"].numpy()
        h[:L] = prompt
        h[L:mp] = 0
        h[mp:mp + L] = range(L)
        h[mp + L:2 * mp] = 0
        h[2 * mp], h[2 * mp + 1], h[2 * mp + 2] = 0, L, L - 1
        ent["dev"].copy_(ent["host"], non_blocking=True)
        ent["evt"] = torch.cuda.Event()
        ent["evt"].record()
        ent["graph"].replay()
        out.copy_(ent["out"])

    def _forward(self, batch: Sequence[Sequence[int]], out: torch.Tensor) -> None:
        lens = [len(p) for p in batch]
        if min(lens) < 1:
            raise RuntimeError("empty prompt")
        if max(lens) > self.max_pos:
            raise RuntimeError(f"prompt of {max(lens)} tokens exceeds max_position_embeddings")
        if len(batch) == 1 and self.use_graphs and (lens[0] + 63) // 64 * 64 <= self.max_pos:
            return self._forward_graph(batch[0], out)
        n, m = len(batch), sum(lens)
        # one pinned staging buffer (persistent, grow-only), one H2D copy:
        # ids | positions | cu_seqlens | last_rows
        need = 2 * m + 2 * n + 1
        slot = self._stage[self._stage_i]
        self._stage_i = (self._stage_i + 1) % len(self._stage)
        if slot[1] is not None:
            slot[1].synchronize()                 # the previous copy out of this buffer is done
        if slot[0] is None or slot[0].numel() < need:
            slot[0] = torch.empty(max(need, 8192), dtype=torch.int32).pin_memory()
        host = slot[0][:need]
        ids, pos = host[:m], host[m:2 * m]
        cu, last = host[2 * m:2 * m + n + 1], host[2 * m + n + 1:]
        import numpy as np
        la = np.asarray(lens, dtype=np.int64)
        ends = np.cumsum(la)
        ids.numpy()[:] = np.fromiter(chain.from_iterable(batch), dtype=np.int32, count=m)
        pos.numpy()[:] = np.arange(m, dtype=np.int64) - np.repeat(ends - la, la)
        cu.numpy()[0] = 0
        cu.numpy()[1:] = ends
        last.numpy()[:] = ends - 1
        d = host.to(self.device, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        ws, ws_bytes = self._workspace(m, n)
        base = d.data_ptr()
        with torch.cuda.device(self.device):
            check(lib.vltr_opt_forward(self._handle, base, base + 4 * m, base + 8 * m,
                                       base + 8 * m + 4 * (n + 1), m, n, max(lens), ws,
                                       ctypes.c_size_t(ws_bytes), out.data_ptr(),
                                       torch._C._cuda_getCurrentRawStream(self.device.index)))
        self._last_inputs = d   # keep alive until the stream has consumed it

    def score_column(self, prompts: Sequence[Sequence[int]]) -> torch.Tensor:
        """``logits[:, 0]`` after compute_logits (opt.py:389-409; for classification heads the
        argmax label as float) as a float32 tensor ON THE DEVICE: nothing here waits for the GPU,
        so a caller can prepare the next batch on the host while this one runs."""
        lg = self.logits(prompts)
        if self.num_labels > 1:
            lg = lg[:, :self.num_labels].argmax(dim=-1, keepdim=True).float()
        return lg[:, 0]

    def scores(self, prompts: Sequence[Sequence[int]], mtype: str = "reg") -> List[float]:
        """What the reference hands to the scheduler: ``logits[:,0].tolist()``."""
        return self.score_column(prompts).tolist()


class B200AuxModel:
    """``scheduler.aux_model``: any object with ``obtain_aux_scores(seq_groups) -> List[float]``
    that calls ``set_aux_model_score`` on each group (aux_llm_engine.py:332-412).

    A burst of new requests (10 000 at t=0 in the synthetic trace) is scored in chunks of
    ``chunk_prompts`` prompts (2048: smaller tokeniser calls lose parallel efficiency,
    profiles/r02_predictor_host_pipeline.txt): the forward of a chunk is only enqueued, so the
    GPU runs it while the host tokenises and packs the next one; the scores are read back once,
    after the last chunk (one device synchronisation per call, as before)."""

    def __init__(self, predictor: OPTPredictor, tokenizer: Callable[[str], List[int]],
                 max_length: int = 2048,
                 batch_tokenizer: Optional[Callable[[List[str]], List[List[int]]]] = None,
                 chunk_prompts: int = 2048):
        self.predictor = predictor
        self.tokenizer = tokenizer
        self.batch_tokenizer = batch_tokenizer
        self.max_length = max_length
        self.chunk_prompts = max(1, int(chunk_prompts))
        self._pool = None

    @staticmethod
    def _batch_encoder(hf) -> Optional[Callable[[List[str]], List[List[int]]]]:
        """Batch encoding with the ids ``hf.encode(text)`` gives one by one, tokenised in parallel
        by the Rust backend.  With no truncation / padding configured on the backend tokenizer its
        own ``encode_batch`` is what ``hf(texts)`` ends up calling
        (PreTrainedTokenizerFast._batch_encode_plus) minus the BatchEncoding bookkeeping
        (attention masks, type ids: ~40 % of the call at 10 000 prompts)."""
        if not getattr(hf, "is_fast", False):
            return None
        raw = getattr(hf, "backend_tokenizer", None)
        if raw is not None and raw.truncation is None and raw.padding is None:
            return lambda texts: [e.ids for e in raw.encode_batch(texts, add_special_tokens=True)]
        return lambda texts: hf(texts, add_special_tokens=True)["input_ids"]

    @classmethod
    def from_pretrained(cls, path: str, tokenizer=None, max_length: int = 2048, **kw):
        pred = OPTPredictor.from_pretrained(path, **kw)
        batch = None
        if tokenizer is None or isinstance(tokenizer, str):
            from transformers import AutoTokenizer
            hf = AutoTokenizer.from_pretrained(tokenizer or path)
            tokenizer = lambda text: hf.encode(text)
            batch = cls._batch_encoder(hf)
        return cls(pred, tokenizer, max_length, batch)

    def _encode(self, texts: List[str]) -> List[List[int]]:
        L = self.max_length
        if self.batch_tokenizer is not None and len(texts) > 1:
            ids = self.batch_tokenizer(texts)
        else:
            ids = [self.tokenizer(t) for t in texts]
        return [(p if isinstance(p, list) else list(p)) if len(p) <= L else list(p[:L])
                for p in ids]

    def obtain_aux_scores(self, seq_groups) -> List[float]:
        texts = []
        for sg in seq_groups:
            assert sg.need_aux_model_score()
            texts.append(sg.seqs_dict[next(iter(sg.seqs_dict))].prompt)
        if not texts:
            return []
        step = self.chunk_prompts
        chunks = [texts[c:c + step] for c in range(0, len(texts), step)]
        parts = []
        if len(chunks) == 1:
            parts.append(self.predictor.score_column(self._encode(chunks[0])))
        else:
            # chunk k+1 is tokenised by a helper thread (the Rust tokenizer releases the GIL)
            # while this thread packs and enqueues the forwards of chunk k; every CUDA call stays
            # on this thread.  Nothing below waits for the device except the staging ring.
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=1,
                                                thread_name_prefix="vltr-aux-tokenize")
            pending = self._pool.submit(self._encode, chunks[0])
            for k in range(len(chunks)):
                prompts = pending.result()
                if k + 1 < len(chunks):
                    pending = self._pool.submit(self._encode, chunks[k + 1])
                parts.append(self.predictor.score_column(prompts))
        scores = (parts[0] if len(parts) == 1 else torch.cat(parts)).tolist()
        for sg, s in zip(seq_groups, scores):
            sg.set_aux_model_score(s)
        return scores
