"""Decode-step hot path: the sequence of ``vllm._C`` calls one engine step makes per decoder layer.

This is the host-side mirror of the call order in the reference's model code, restricted to the
ops in scope:

    LlamaDecoderLayer.forward   vllm/model_executor/models/llama.py:215-238
      input_layernorm           -> ops.rms_norm (first layer) / ops.fused_add_rms_norm
      LlamaAttention.forward    llama.py:156-169
        qkv_proj                   (cuBLAS GEMM - library code, not part of this path)
        rotary_emb              -> ops.rotary_embedding          layers/rotary_embedding.py:143-161
        Attention.forward       -> cache_ops.reshape_and_cache   attention/backends/flash_attn.py:192
                                -> ops.paged_attention_v1|v2     attention/ops/paged_attn.py:83-161
        o_proj                     (cuBLAS GEMM [+ all-reduce when TP>1])
      post_attention_layernorm  -> ops.fused_add_rms_norm
      LlamaMLP.forward          llama.py:75-79
        gate_up_proj               (cuBLAS GEMM)
        act_fn                  -> ops.silu_and_mul              layers/activation.py:31-36
        down_proj                  (cuBLAS GEMM [+ all-reduce when TP>1])
    LlamaModel.forward          llama.py:271-303: final norm -> ops.fused_add_rms_norm

The GEMM outputs (qkv, o_proj, gate_up, down_proj) are stand-in buffers owned by the caller; the
ops read and write them in place exactly as the reference layers do.  Everything goes through a
``vllm._C``-shaped module: ``vllm_ltr_b200._C`` by default, or the reference's own builds
(oracle/_ref/ref_cuda_C, ref_cpu_C) for the baseline arms.

BENCH / TEST DRIVER (tools/), not product code: the product is the extension surface itself; in
the engine this call order is the reference's model code.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch



def _default_C():
    from vllm_ltr_b200 import _C      # loads libvllm_ltr_b200.so; not needed by the reference arms
    return _C


_PARTITION_SIZE = 512  # vllm/attention/ops/paged_attn.py:10


@dataclass
class ModelShape:
    """Per-GPU shapes of one decoder layer (after tensor-parallel sharding)."""
    name: str
    hidden: int
    layers: int
    q_heads: int
    kv_heads: int
    head_size: int
    inter: int
    dtype: torch.dtype
    kv_cache_dtype: str = "auto"
    block_size: int = 16
    rope_theta: float = 10000.0
    max_pos: int = 8192
    eps: float = 1e-5

    def shard(self, tp: int) -> "ModelShape":
        """Megatron column/row split used by the reference (llama.py:96-111): q heads and the
        MLP width are divided, kv heads are divided or replicated."""
        if tp == 1:
            return self
        assert self.q_heads % tp == 0 and self.inter % tp == 0
        return ModelShape(f"{self.name}-tp{tp}", self.hidden, self.layers, self.q_heads // tp,
                          max(1, self.kv_heads // tp), self.head_size, self.inter // tp,
                          self.dtype, self.kv_cache_dtype, self.block_size, self.rope_theta,
                          self.max_pos, self.eps)


SHAPES = {
    "llama-2-7b": ModelShape("llama-2-7b", 4096, 32, 32, 32, 128, 11008, torch.float16),
    "llama-3-8b": ModelShape("llama-3-8b", 4096, 32, 32, 8, 128, 14336, torch.bfloat16,
                             rope_theta=500000.0),
    "llama-3-8b-fp8kv": ModelShape("llama-3-8b-fp8kv", 4096, 32, 32, 8, 128, 14336,
                                   torch.bfloat16, "fp8", rope_theta=500000.0),
    "llama-2-70b": ModelShape("llama-2-70b", 8192, 80, 64, 8, 128, 28672, torch.float16),
}


def cos_sin_cache(shape: ModelShape, device) -> torch.Tensor:
    """RotaryEmbedding._compute_cos_sin_cache, layers/rotary_embedding.py:70-93."""
    d = shape.head_size
    inv_freq = 1.0 / (shape.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float) / d))
    t = torch.arange(shape.max_pos, dtype=torch.float)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(shape.dtype).to(device)


def forward_decode(query, key_cache, value_cache, block_tables, context_lens, max_context_len,
                   kv_cache_dtype, num_kv_heads, scale, alibi_slopes, kv_scale, out=None,
                   scratch=None, C=None):
    """PagedAttention.forward_decode (attention/ops/paged_attn.py:83-161), including the
    reference's v1/v2 choice.  ``out`` / ``scratch`` let a caller reuse buffers across steps."""
    C = _default_C() if C is None else C
    output = torch.empty_like(query) if out is None else out
    block_size = value_cache.shape[3]
    num_seqs, num_heads, head_size = query.shape
    parts = (max_context_len + _PARTITION_SIZE - 1) // _PARTITION_SIZE
    use_v1 = max_context_len <= 8192 and (parts == 1 or num_seqs * num_heads > 512)
    if use_v1:
        C.ops.paged_attention_v1(output, query, key_cache, value_cache, num_kv_heads, scale,
                                  block_tables, context_lens, block_size, max_context_len,
                                  alibi_slopes, kv_cache_dtype, kv_scale)
    else:
        if scratch is None:
            tmp = torch.empty(num_seqs, num_heads, parts, head_size, dtype=output.dtype,
                              device=output.device)
            es = torch.empty(num_seqs, num_heads, parts, dtype=torch.float32,
                             device=output.device)
            ml = torch.empty_like(es)
        else:
            tmp, es, ml = scratch
        C.ops.paged_attention_v2(output, es, ml, tmp, query, key_cache, value_cache,
                                  num_kv_heads, scale, block_tables, context_lens, block_size,
                                  max_context_len, alibi_slopes, kv_cache_dtype, kv_scale)
    return output


class DecodeHotPath:
    """State of one model replica's decode step: paged KV pool, norm weights, rope cache and the
    stand-in GEMM output buffers; ``step()`` issues the per-layer op sequence."""

    def __init__(self, shape: ModelShape, num_blocks: int, max_tokens: int, device, seed: int = 0,
                 layers: Optional[int] = None, C=None):
        # C: the extension module to drive.  Always vllm_ltr_b200._C in the product; bench.py's
        # reference arm passes the reference's own CPU build of vllm._C (same surface) instead.
        self.C = _default_C() if C is None else C
        # rope + cache append as one launch: only with the product's own extension
        self.fuse_rope_cache = C is None or getattr(C, "__name__", "").startswith("vllm_ltr_b200")
        self.shape = s = shape
        self.device = device
        self.layers = s.layers if layers is None else layers
        self.num_blocks = num_blocks
        g = torch.Generator(device=device).manual_seed(seed)
        es = 1 if s.kv_cache_dtype == "fp8" else torch.tensor([], dtype=s.dtype).element_size()
        self.x = 16 // es
        cache_t = torch.uint8 if s.kv_cache_dtype == "fp8" else s.dtype
        scale = s.head_size ** -0.5
        # one [2, num_blocks, block_size*kv_heads*head_size] tensor per layer, re-viewed as the
        # reference does (paged_attn.py:42-61)
        self.kv_caches: List[torch.Tensor] = []
        self.key_caches: List[torch.Tensor] = []
        self.value_caches: List[torch.Tensor] = []
        for _ in range(self.layers):
            n = s.block_size * s.kv_heads * s.head_size
            if s.kv_cache_dtype == "fp8":
                kv = torch.randint(0, 0x38, (2, num_blocks, n), generator=g, device=device,
                                   dtype=torch.uint8)          # finite e5m2 codes below 1.0
            else:
                kv = ((torch.rand(2, num_blocks, n, generator=g, device=device) * 2 - 1)
                      * scale).to(cache_t)
            self.kv_caches.append(kv)
            self.key_caches.append(kv[0].view(num_blocks, s.kv_heads, s.head_size // self.x,
                                              s.block_size, self.x))
            self.value_caches.append(kv[1].view(num_blocks, s.kv_heads, s.head_size,
                                                s.block_size))
        self.norm_w = [(torch.randn(s.hidden, generator=g, device=device) * 0.1 + 1).to(s.dtype)
                       for _ in range(2 * self.layers + 1)]
        self.cos_sin = cos_sin_cache(s, device)
        T = max_tokens
        qkv_w = (s.q_heads + 2 * s.kv_heads) * s.head_size
        rnd = lambda *sh: (torch.randn(*sh, generator=g, device=device) * 0.5).to(s.dtype)
        # stand-ins for the cuBLAS GEMM outputs of one layer (reused by every layer)
        self.qkv = rnd(T, qkv_w)
        self.o_proj_out = rnd(T, s.hidden)
        self.gate_up = rnd(T, 2 * s.inter)
        self.down_out = rnd(T, s.hidden)
        self.attn_out = torch.empty(T, s.q_heads * s.head_size, dtype=s.dtype, device=device)
        self.act_out = torch.empty(T, s.inter, dtype=s.dtype, device=device)
        self.hidden = rnd(T, s.hidden)
        self.residual = torch.empty_like(self.hidden)
        self.normed = torch.empty_like(self.hidden)
        self.scale = float(scale)
        self.attn_events = None   # optional [(start, end)] CUDA events around each attention call
        self.gemm_w = None        # optional per-layer GEMM weights (enable_gemms)

    def _rope_and_cache(self, positions, q, k, v, kc, vc, slot_mapping, T) -> None:
        """rotary_emb + the cache append of Attention.forward (llama.py:159, flash_attn.py:192):
        one fused launch behind the attention seam (vllm_ltr_b200.attention_backend) when this
        path drives the product's kernels, the reference's two ops otherwise (``fuse_rope_cache``
        False: the reference extension modules, or an A/B run)."""
        s = self.shape
        if self.fuse_rope_cache:
            from vllm_ltr_b200.attention_backend import rotary_embedding_and_cache
            if rotary_embedding_and_cache(positions, q, k, v, kc, vc, slot_mapping, s.head_size,
                                          self.cos_sin, True, s.kv_cache_dtype):
                return
        self.C.ops.rotary_embedding(positions, q, k, s.head_size, self.cos_sin, True)
        self.C.cache_ops.reshape_and_cache(k.view(T, s.kv_heads, s.head_size),
                                           v.view(T, s.kv_heads, s.head_size), kc, vc,
                                           slot_mapping, s.kv_cache_dtype, 1.0)

    def enable_gemms(self, seed: int = 1) -> None:
        """Context for the hot-path numbers: also run the four dense projections of every layer
        (qkv_proj, o_proj, gate_up_proj, down_proj - library code, cuBLAS via torch.matmul) with
        seeded random weights of the model's shapes, writing the buffers the ops consume.  The
        step then is the whole decoder stack of a decode step (no embedding / lm_head / sampler);
        weights are per layer, so their HBM traffic is real.  Single GPU only."""
        s = self.shape
        g = torch.Generator(device=self.device).manual_seed(seed)
        qkv_w = (s.q_heads + 2 * s.kv_heads) * s.head_size

        def w(n, k):
            return (torch.randn(n, k, generator=g, device=self.device) * (k ** -0.5)).to(s.dtype)

        self.gemm_w = [(w(qkv_w, s.hidden), w(s.hidden, s.q_heads * s.head_size),
                        w(2 * s.inter, s.hidden), w(s.hidden, s.inter)) for _ in range(self.layers)]

    def capture(self, hidden_in, positions, slot_mapping, block_tables, context_lens,
                max_context_len: int):
        """Record one decode step into a CUDA graph over the given (static) input tensors and
        return ``(graph, output)``: the caller refreshes the inputs in place and calls
        ``graph.replay()``.  ``max_context_len`` is baked into the recorded launches (it sizes the
        attention partitions), so a graph is valid for contexts up to that bound - the engine
        keys its graphs the same way (model_runner.py:896-970 pads to the captured batch size and
        passes the model's max context length).  The first, uncaptured step lets the library
        grow its workspaces; nothing allocates while recording.  With the fused tensor-parallel
        exchange attached, every rank must capture and replay in lockstep (each replay is one
        set of exchanges) and the communicator must count its epochs on the device."""
        comm = getattr(self, "comm", None)
        if comm is not None and not comm.device_epoch:
            raise RuntimeError("capture(): FusedTPComm(device_epoch=False) passes the barrier "
                               "epoch as a launch argument and cannot be replayed")
        assert self.attn_events is None
        args = (hidden_in, positions, slot_mapping, block_tables, context_lens, max_context_len)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.step(*args)
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.step(*args)
        return graph, out

    def attention_only(self, block_tables, context_lens, max_context_len: int, T: int) -> None:
        """The paged-attention calls of one step and nothing else: one launch per layer, each
        over that layer's own KV pool, back to back on the current stream (bench.py times this
        between one pair of events to get the kernel's in-stream launch duration)."""
        s = self.shape
        qsz = s.q_heads * s.head_size
        q = self.qkv[:T, :qsz].view(T, s.q_heads, s.head_size)
        out = self.attn_out[:T].view(T, s.q_heads, s.head_size)
        for layer in range(self.layers):
            forward_decode(q, self.key_caches[layer], self.value_caches[layer], block_tables,
                           context_lens, max_context_len, s.kv_cache_dtype, s.kv_heads, self.scale,
                           None, 1.0, out=out, C=self.C)

    def attach_fused_tp(self, comm) -> None:
        """Tensor-parallel mode with the fused exchange (vllm_ltr_b200.parallel.FusedTPComm): the
        row-parallel GEMM outputs, the residual and the normed activations live in the
        communicator's symmetric buffers; every ``all_reduce + fused_add_rms_norm`` pair of the
        layer becomes one ``allreduce_add_rms_norm`` call."""
        self.comm = comm
        comm.x[0][:self.o_proj_out.shape[0]].copy_(self.o_proj_out)
        comm.x[1][:self.down_out.shape[0]].copy_(self.down_out)

    def _step_fused_tp(self, hidden_in, positions, slot_mapping, block_tables, context_lens,
                       max_context_len) -> torch.Tensor:
        s, comm = self.shape, self.comm
        ops, cache_ops = self.C.ops, self.C.cache_ops
        T = hidden_in.shape[0]
        qsz, kvsz = s.q_heads * s.head_size, s.kv_heads * s.head_size
        q, k, v = self.qkv[:T].split([qsz, kvsz, kvsz], dim=-1)
        attn_out = self.attn_out[:T]
        for layer in range(self.layers):
            if layer == 0:
                # embedding output is replicated on every rank: plain local norm
                comm.resid[:T].copy_(hidden_in)
                ops.rms_norm(comm.out[:T], comm.resid[:T], self.norm_w[0], s.eps)
            else:
                # previous down_proj exchange + this layer's input_layernorm
                comm.allreduce_add_rms_norm(1, self.norm_w[2 * layer], s.eps, T)
            kc, vc = self.key_caches[layer], self.value_caches[layer]
            self._rope_and_cache(positions, q, k, v, kc, vc, slot_mapping, T)
            ev = self.attn_events
            if ev is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            forward_decode(q.view(T, s.q_heads, s.head_size), kc, vc, block_tables, context_lens,
                           max_context_len, s.kv_cache_dtype, s.kv_heads, self.scale, None, 1.0,
                           out=attn_out.view(T, s.q_heads, s.head_size), C=self.C)
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
            # o_proj exchange + post_attention_layernorm
            comm.allreduce_add_rms_norm(0, self.norm_w[2 * layer + 1], s.eps, T)
            ops.silu_and_mul(self.act_out[:T], self.gate_up[:T])
        # last down_proj exchange + final norm
        return comm.allreduce_add_rms_norm(1, self.norm_w[2 * self.layers], s.eps, T)

    def step(self, hidden_in: torch.Tensor, positions: torch.Tensor, slot_mapping: torch.Tensor,
             block_tables: torch.Tensor, context_lens: torch.Tensor, max_context_len: int,
             all_reduce=None) -> torch.Tensor:
        """One decode step for ``T = hidden_in.shape[0]`` sequences (one new token each).
        Returns the final-norm output [T, hidden] (what compute_logits consumes)."""
        if getattr(self, "comm", None) is not None:
            return self._step_fused_tp(hidden_in, positions, slot_mapping, block_tables,
                                       context_lens, max_context_len)
        s = self.shape
        ops, cache_ops = self.C.ops, self.C.cache_ops
        T = hidden_in.shape[0]
        qsz, kvsz = s.q_heads * s.head_size, s.kv_heads * s.head_size
        qkv = self.qkv[:T]
        q, k, v = qkv.split([qsz, kvsz, kvsz], dim=-1)
        hidden = self.hidden[:T]
        hidden.copy_(hidden_in)
        residual = self.residual[:T]
        attn_out = self.attn_out[:T]
        for layer in range(self.layers):
            # input_layernorm
            if layer == 0:
                residual.copy_(hidden)
                ops.rms_norm(self.normed[:T], hidden, self.norm_w[0], s.eps)
            else:
                ops.fused_add_rms_norm(hidden, residual, self.norm_w[2 * layer], s.eps)
            # [qkv_proj GEMM writes qkv]
            gw = self.gemm_w[layer] if self.gemm_w is not None else None
            if gw is not None:
                torch.matmul(self.normed[:T] if layer == 0 else hidden, gw[0].t(), out=qkv)
            kc, vc = self.key_caches[layer], self.value_caches[layer]
            self._rope_and_cache(positions, q, k, v, kc, vc, slot_mapping, T)
            ev = self.attn_events
            if ev is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            forward_decode(q.view(T, s.q_heads, s.head_size), kc, vc, block_tables, context_lens,
                           max_context_len, s.kv_cache_dtype, s.kv_heads, self.scale, None, 1.0,
                           out=attn_out.view(T, s.q_heads, s.head_size), C=self.C)
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
            # [o_proj GEMM reads attn_out, writes o_proj_out; row-parallel => all-reduce]
            hidden = self.o_proj_out[:T]
            if gw is not None:
                torch.matmul(attn_out, gw[1].t(), out=hidden)
            if all_reduce is not None:
                all_reduce(hidden)
            ops.fused_add_rms_norm(hidden, residual, self.norm_w[2 * layer + 1], s.eps)
            # [gate_up_proj GEMM writes gate_up]
            if gw is not None:
                torch.matmul(hidden, gw[2].t(), out=self.gate_up[:T])
            ops.silu_and_mul(self.act_out[:T], self.gate_up[:T])
            # [down_proj GEMM reads act_out, writes down_out; row-parallel => all-reduce]
            hidden = self.down_out[:T]
            if gw is not None:
                torch.matmul(self.act_out[:T], gw[3].t(), out=hidden)
            if all_reduce is not None:
                all_reduce(hidden)
        ops.fused_add_rms_norm(hidden, residual, self.norm_w[2 * self.layers], s.eps)
        return hidden
