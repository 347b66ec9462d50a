"""The reference's second seam for this path: an ``AttentionBackend`` / ``AttentionImpl`` pair
(vllm/attention/backends/abstract.py:8-124) whose ``forward`` issues the hot-path ops of
``libvllm_ltr_b200`` directly - cache append, prefill attention over the paged cache, decode
attention - instead of going through ``PagedAttention``'s wrappers and the Triton prefix kernel
(vllm/attention/backends/flash_attn.py:154-268, vllm/attention/ops/paged_attn.py:63-195).

The reference picks its backend from an enum at model-construction time
(vllm/attention/selector.py:16-87 - adding a member means editing that file), so the drop-in
route is ``install(model)``: every ``vllm.attention.layer.Attention`` module of an already built
model gets a ``B200AttentionImpl`` with the arguments its own impl was built with
(attention/layer.py:34-37).  Metadata, KV-cache shape and the swap / copy entry points are the
reference's FlashAttention backend's, unchanged (same wire format, model_runner.py:740-800).

This is also where a fused cache-append + attention kernel plugs in: ``forward`` is the one place
that sees key / value, the slot mapping and the block tables of a step together.
"""
import ctypes
from typing import Dict, List, Optional, Tuple, Type

import torch

from ._lib import check, declare, lib

_PARTITION_SIZE = 512            # vllm/attention/ops/paged_attn.py:10
_P = ctypes.c_void_p
declare("vltr_rotary_embedding_and_cache", ctypes.c_int,
        [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
         ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
         ctypes.c_int, ctypes.c_int, ctypes.c_int, _P])
_DTYPE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_ERR_UNSUPPORTED = 3             # include/vllm_ltr_b200.h


def rotary_embedding_and_cache(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                               value: torch.Tensor, key_cache: torch.Tensor,
                               value_cache: torch.Tensor, slot_mapping: torch.Tensor,
                               head_size: int, cos_sin_cache: torch.Tensor, is_neox: bool,
                               kv_cache_dtype: str = "auto") -> bool:
    """``ops.rotary_embedding(positions, query, key, head_size, cos_sin_cache, is_neox)`` followed
    by ``cache_ops.reshape_and_cache(key, value, key_cache, value_cache, slot_mapping, ...)`` in
    one launch (csrc/rope_cache.cu).  query [T, Hq*D] / key, value [T, Hkv*D] (or 3-D views) with
    unit element stride inside a row; query and key are rotated in place.  Returns False - and
    touches nothing - when the shapes are outside the fused form (the caller issues the two ops);
    other errors raise."""
    if not query.is_cuda:
        raise RuntimeError("vllm_ltr_b200: tensors must be on a CUDA device (no CPU path)")
    code = _DTYPE.get(query.dtype)
    if code is None or key.dtype != query.dtype or value.dtype != query.dtype or \
            cos_sin_cache.dtype != query.dtype:
        return False
    if kv_cache_dtype not in ("auto", "fp8"):
        raise RuntimeError(f"Unsupported data type of kv cache: {kv_cache_dtype}")
    fp8 = 1 if kv_cache_dtype == "fp8" else 0
    if positions.dtype != torch.int64 or slot_mapping.dtype != torch.int64 or \
            not positions.is_contiguous() or not slot_mapping.is_contiguous():
        return False
    T = query.shape[0]
    Hq = query.numel() // (T * head_size) if T else 0
    Hkv = key.numel() // (T * head_size) if T else 0
    if T == 0:
        return True
    if positions.numel() != T or slot_mapping.numel() != T or \
            query.stride(-1) != 1 or key.stride(-1) != 1 or value.stride(-1) != 1:
        return False
    rot_dim = cos_sin_cache.shape[1]
    with torch.cuda.device(query.device):
        rc = lib.vltr_rotary_embedding_and_cache(
            positions.data_ptr(), query.data_ptr(), key.data_ptr(), value.data_ptr(),
            key_cache.data_ptr(), value_cache.data_ptr(), slot_mapping.data_ptr(),
            cos_sin_cache.data_ptr(), T, Hq, Hkv, head_size, rot_dim, 1 if is_neox else 0,
            query.stride(0), key.stride(0), value.stride(0), key_cache.size(3), key_cache.size(4),
            code, fp8, torch._C._cuda_getCurrentRawStream(query.device.index))
    if rc == _ERR_UNSUPPORTED:
        return False
    check(rc)
    return True

_classes = None


def backend_classes():
    """(B200AttentionBackend, B200AttentionImpl), created on first use: they subclass the
    reference's abstract classes, which are importable only where the engine is."""
    global _classes
    if _classes is not None:
        return _classes
    from vllm.attention.backends.abstract import AttentionBackend, AttentionImpl
    from vllm.attention.backends.flash_attn import FlashAttentionMetadata
    from vllm.attention.ops.paged_attn import PagedAttention

    from . import _C
    from .prefix_prefill import context_attention_fwd

    class B200AttentionImpl(AttentionImpl):

        def __init__(self, num_heads: int, head_size: int, scale: float,
                     num_kv_heads: Optional[int] = None,
                     alibi_slopes: Optional[List[float]] = None,
                     sliding_window: Optional[int] = None, reference_impl=None) -> None:
            if sliding_window is not None and sliding_window != (-1, -1):
                raise NotImplementedError("B200AttentionImpl: sliding-window attention is not "
                                          "on this path (none of the LTR configs uses it)")
            self.num_heads = num_heads
            self.head_size = head_size
            self.scale = float(scale)
            self.num_kv_heads = num_heads if num_kv_heads is None else num_kv_heads
            if alibi_slopes is not None and not torch.is_tensor(alibi_slopes):
                alibi_slopes = torch.tensor(alibi_slopes, dtype=torch.float32)
            self.alibi_slopes = alibi_slopes
            assert self.num_heads % self.num_kv_heads == 0
            if head_size not in PagedAttention.get_supported_head_sizes():
                raise ValueError(f"Head size {head_size} is not supported by PagedAttention.")
            # only for the memory-profiling run, which has no KV cache (model_runner.py:863-869)
            self.reference_impl = reference_impl
            # (positions, RotaryEmbedding) left by a DeferredRotary in front of this layer: the
            # rotation is applied here, fused with the cache append
            self._pending_rope = None

        def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                    kv_cache: Optional[torch.Tensor], attn_metadata, kv_scale: float) -> torch.Tensor:
            rope, self._pending_rope = self._pending_rope, None
            if kv_cache is None:
                if self.reference_impl is None:
                    raise RuntimeError("B200AttentionImpl needs a KV cache (the cache-less "
                                       "profiling run belongs to the reference impl)")
                if rope is not None:
                    rope[1](rope[0], query, key)
                return self.reference_impl.forward(query, key, value, kv_cache, attn_metadata,
                                                   kv_scale)
            num_tokens, hidden = query.shape
            H, Hkv, D = self.num_heads, self.num_kv_heads, self.head_size
            q = query.view(-1, H, D)
            k = key.view(-1, Hkv, D)
            v = value.view(-1, Hkv, D)
            key_cache, value_cache = PagedAttention.split_kv_cache(kv_cache, Hkv, D)
            kvd = attn_metadata.kv_cache_dtype
            # a6 + a3: rotate this step's queries / keys (when a DeferredRotary handed the
            # rotation over) and append keys / values to the cache (flash_attn.py:185-196) - one
            # launch; the two reference ops when the shapes are outside the fused form
            slots = attn_metadata.slot_mapping.flatten()
            fused = False
            if rope is not None:
                positions, rot = rope
                rot.cos_sin_cache = rot.cos_sin_cache.to(positions.device)   # rotary_embedding.py:143
                fused = rotary_embedding_and_cache(positions, q, k, v, key_cache, value_cache,
                                                   slots, D, rot.cos_sin_cache,
                                                   rot.is_neox_style, kvd)
                if not fused:
                    rot(positions, query, key)
            if not fused:
                _C.cache_ops.reshape_and_cache(k, v, key_cache, value_cache, slots, kvd, kv_scale)
            npre = attn_metadata.num_prefill_tokens
            ndec = attn_metadata.num_decode_tokens
            assert k.shape[0] == npre + ndec
            out = torch.empty_like(q)
            if self.alibi_slopes is not None and self.alibi_slopes.device != q.device:
                self.alibi_slopes = self.alibi_slopes.to(q.device)
            pm = attn_metadata.prefill_metadata
            if pm is not None:
                # f1: every prefill chunk, first or continued, through the paged cache (the chunk
                # itself was appended above): flash_attn.py:217-252 without flash-attn / Triton
                if pm.block_tables.numel() == 0:
                    raise RuntimeError("B200AttentionImpl: prefill without block tables (the "
                                       "engine always passes them under chunked prefill)")
                context_attention_fwd(q[:npre], k[:npre], v[:npre], out[:npre], key_cache,
                                      value_cache, pm.block_tables, pm.subquery_start_loc,
                                      pm.prompt_lens_tensor, pm.context_lens, pm.max_subquery_len,
                                      self.alibi_slopes, cache_has_new_tokens=True)
            dm = attn_metadata.decode_metadata
            if dm is not None:
                # a1 / a2 with the reference's choice between them (paged_attn.py:101-110)
                dq, do = q[npre:], out[npre:]
                mx = dm.max_context_len
                parts = (mx + _PARTITION_SIZE - 1) // _PARTITION_SIZE
                bs = value_cache.shape[3]
                if mx <= 8192 and (parts == 1 or dq.shape[0] * H > 512):
                    _C.ops.paged_attention_v1(do, dq, key_cache, value_cache, Hkv, self.scale,
                                              dm.block_tables, dm.context_lens, bs, mx,
                                              self.alibi_slopes, kvd, kv_scale)
                else:
                    n = dq.shape[0]
                    tmp = torch.empty(n, H, parts, D, dtype=do.dtype, device=do.device)
                    es = torch.empty(n, H, parts, dtype=torch.float32, device=do.device)
                    ml = torch.empty_like(es)
                    _C.ops.paged_attention_v2(do, es, ml, tmp, dq, key_cache, value_cache, Hkv,
                                              self.scale, dm.block_tables, dm.context_lens, bs, mx,
                                              self.alibi_slopes, kvd, kv_scale)
            return out.view(-1, H * D)

    class B200AttentionBackend(AttentionBackend):

        @staticmethod
        def get_impl_cls() -> Type["B200AttentionImpl"]:
            return B200AttentionImpl

        @staticmethod
        def make_metadata(*args, **kwargs):
            return FlashAttentionMetadata(*args, **kwargs)

        @staticmethod
        def get_kv_cache_shape(num_blocks: int, block_size: int, num_kv_heads: int,
                               head_size: int) -> Tuple[int, ...]:
            return PagedAttention.get_kv_cache_shape(num_blocks, block_size, num_kv_heads,
                                                     head_size)

        @staticmethod
        def swap_blocks(src_kv_cache: torch.Tensor, dst_kv_cache: torch.Tensor,
                        src_to_dst: Dict[int, int]) -> None:
            _C.cache_ops.swap_blocks(src_kv_cache[0], dst_kv_cache[0], src_to_dst)
            _C.cache_ops.swap_blocks(src_kv_cache[1], dst_kv_cache[1], src_to_dst)

        @staticmethod
        def copy_blocks(kv_caches: List[torch.Tensor],
                        src_to_dists: Dict[int, List[int]]) -> None:
            _C.cache_ops.copy_blocks([c[0] for c in kv_caches], [c[1] for c in kv_caches],
                                     src_to_dists)

    _classes = (B200AttentionBackend, B200AttentionImpl)
    return _classes


class DeferredRotary(torch.nn.Module):
    """Stands where a decoder layer keeps its ``rotary_emb`` (llama.py:139-147, called at :159
    right before ``self.attn``): instead of launching ``ops.rotary_embedding`` it hands
    (positions, the real module) to the layer's ``B200AttentionImpl``, whose ``forward`` - the
    very next call - rotates queries and keys in place while appending to the KV cache
    (csrc/rope_cache.cu).  Calls the fused kernel cannot serve (per-token offsets of the batched
    LoRA form) go straight to the real module."""

    def __init__(self, inner: torch.nn.Module, impl) -> None:
        super().__init__()
        self.inner = inner
        self._impl = [impl]              # in a list: not a submodule, not in state_dict

    def forward(self, positions, query, key, offsets=None):
        if offsets is not None:
            return self.inner(positions, query, key, offsets)
        self._impl[0]._pending_rope = (positions, self.inner)
        return query, key


def install(model: torch.nn.Module, fuse_rope: bool = True) -> int:
    """Give every ``Attention`` layer of a built model a ``B200AttentionImpl`` and, with
    ``fuse_rope``, defer the rotary embedding of its decoder layer into it.  Returns the number
    of layers switched."""
    from vllm.attention.layer import Attention
    backend, impl_cls = backend_classes()
    n = 0
    if fuse_rope:
        from vllm.model_executor.layers.rotary_embedding import RotaryEmbedding
        for parent in model.modules():
            rot = getattr(parent, "rotary_emb", None)
            att = getattr(parent, "attn", None)
            if isinstance(att, Attention) and isinstance(rot, RotaryEmbedding) and \
                    type(rot).forward is RotaryEmbedding.forward and rot.is_neox_style and \
                    rot.rotary_dim == rot.head_size:
                if not isinstance(att.impl, impl_cls):
                    old = att.impl
                    sw = getattr(old, "sliding_window", None)
                    att.impl = impl_cls(old.num_heads, old.head_size, old.scale, old.num_kv_heads,
                                        old.alibi_slopes, None if sw in (None, (-1, -1)) else sw,
                                        reference_impl=old)
                    att.backend = backend
                    n += 1
                parent.rotary_emb = DeferredRotary(rot, att.impl)
    for mod in model.modules():
        if isinstance(mod, Attention) and not isinstance(mod.impl, impl_cls):
            old = mod.impl
            sw = getattr(old, "sliding_window", None)
            mod.impl = impl_cls(old.num_heads, old.head_size, old.scale, old.num_kv_heads,
                                old.alibi_slopes, None if sw in (None, (-1, -1)) else sw,
                                reference_impl=old)
            mod.backend = backend
            n += 1
    return n
