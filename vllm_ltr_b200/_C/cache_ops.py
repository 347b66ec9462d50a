"""``vllm._C.cache_ops`` — swap_blocks, copy_blocks, reshape_and_cache, convert_fp8
(csrc/pybind.cpp:77-94, declarations csrc/cache.h:8-30)."""
import ctypes
from typing import Dict, List

import torch

from ._common import DTYPE_CODE, U8, check, device_of, float_code, lib, require_cuda, stream_of


def _pairs(flat: List[int]):
    n = len(flat)
    return (ctypes.c_int64 * n)(*flat)


def swap_blocks(src: torch.Tensor, dst: torch.Tensor, block_mapping: Dict[int, int]) -> None:
    """csrc/cache_kernels.cu:23-62.  src/dst: one layer's K or V cache on GPU or CPU."""
    if src.is_cuda and dst.is_cuda:
        if src.device.index != dst.device.index:
            raise RuntimeError("src and dst must be on the same GPU")
        kind, dev_t = 0, src
    elif src.is_cuda and dst.device.type == "cpu":
        kind, dev_t = 1, src
    elif src.device.type == "cpu" and dst.is_cuda:
        kind, dev_t = 2, dst
    else:
        raise RuntimeError("Invalid device combination")
    block_bytes = src.element_size() * src[0].numel()
    flat: List[int] = []
    for s, d in sorted(block_mapping.items()):  # std::map iteration order
        flat.append(s)
        flat.append(d)
    with device_of(dev_t):
        check(lib.vltr_swap_blocks(src.data_ptr(), dst.data_ptr(), kind, _pairs(flat),
                                   len(flat) // 2, block_bytes, stream_of(dev_t)))


def swap_blocks_layers(src_kv_caches: List[torch.Tensor], dst_kv_caches: List[torch.Tensor],
                       block_mapping: Dict[int, int]) -> None:
    """Every layer of CacheEngine.swap_in / swap_out (vllm/worker/cache_engine.py:92-100) in one
    call: ``*_kv_caches[i]`` is layer i's ``[2, num_blocks, ...]`` cache on the GPU or in (pinned)
    host memory; the reference loops ``PagedAttention.swap_blocks`` over the layers
    (paged_attn.py:195-206).  Not part of the reference's ``vllm._C`` surface."""
    if len(src_kv_caches) != len(dst_kv_caches):
        raise RuntimeError("src and dst must have the same number of layers")
    if not src_kv_caches:
        return
    s0, d0 = src_kv_caches[0], dst_kv_caches[0]
    if s0.is_cuda and d0.is_cuda:
        kind, dev_t = 0, s0
    elif s0.is_cuda and d0.device.type == "cpu":
        kind, dev_t = 1, s0
    elif s0.device.type == "cpu" and d0.is_cuda:
        kind, dev_t = 2, d0
    else:
        raise RuntimeError("Invalid device combination")
    bufs_s, bufs_d = [], []
    for s, d in zip(src_kv_caches, dst_kv_caches):
        if s.device != s0.device or d.device != d0.device or s.shape[0] != 2 or d.shape[0] != 2:
            raise RuntimeError("every layer must be a [2, num_blocks, ...] cache on the same devices")
        bufs_s += [s[0].data_ptr(), s[1].data_ptr()]
        bufs_d += [d[0].data_ptr(), d[1].data_ptr()]
    block_bytes = s0.element_size() * s0[0][0].numel()
    flat: List[int] = []
    for s, d in sorted(block_mapping.items()):
        flat.append(s)
        flat.append(d)
    n = len(bufs_s)
    with device_of(dev_t):
        check(lib.vltr_swap_blocks_layers((ctypes.c_void_p * n)(*bufs_s), (ctypes.c_void_p * n)(*bufs_d),
                                          n, kind, _pairs(flat), len(flat) // 2, block_bytes,
                                          stream_of(dev_t)))


def install_cache_engine(cache_engine) -> None:
    """Re-bind ``swap_in`` / ``swap_out`` of a reference ``CacheEngine``
    (vllm/worker/cache_engine.py:92-100) to the one-launch form."""
    cache_engine.swap_in = lambda m: swap_blocks_layers(cache_engine.cpu_cache, cache_engine.gpu_cache, m)
    cache_engine.swap_out = lambda m: swap_blocks_layers(cache_engine.gpu_cache, cache_engine.cpu_cache, m)


def copy_blocks(key_caches: List[torch.Tensor], value_caches: List[torch.Tensor],
                block_mapping: Dict[int, List[int]]) -> None:
    """csrc/cache_kernels.cu:97-152."""
    num_layers = len(key_caches)
    if num_layers != len(value_caches):
        raise RuntimeError("key_caches and value_caches must have the same length")
    if num_layers == 0:
        return
    require_cuda(key_caches[0], "key_caches")
    flat: List[int] = []
    for s, dsts in sorted(block_mapping.items()):
        for d in dsts:
            flat.append(s)
            flat.append(d)
    kptrs = (ctypes.c_void_p * num_layers)(*[t.data_ptr() for t in key_caches])
    vptrs = (ctypes.c_void_p * num_layers)(*[t.data_ptr() for t in value_caches])
    k0 = key_caches[0]
    block_bytes = k0[0].numel() * k0.element_size()
    with device_of(k0):
        check(lib.vltr_copy_blocks(kptrs, vptrs, num_layers, _pairs(flat), len(flat) // 2,
                                   block_bytes, stream_of(k0)))


def reshape_and_cache(key: torch.Tensor, value: torch.Tensor, key_cache: torch.Tensor,
                      value_cache: torch.Tensor, slot_mapping: torch.Tensor, kv_cache_dtype: str,
                      kv_scale: float) -> None:
    """csrc/cache_kernels.cu:235-276."""
    require_cuda(key, "key")
    code = float_code(key, "key")
    if kv_cache_dtype == "auto":
        fp8 = 0
    elif kv_cache_dtype == "fp8":
        fp8 = 1
    else:
        raise RuntimeError(f"Unsupported data type of kv cache: {kv_cache_dtype}")
    if slot_mapping.dtype != torch.int64:
        raise RuntimeError("slot_mapping must be int64")
    if (key_cache.dtype != torch.uint8) if fp8 else (key_cache.dtype != key.dtype):
        raise RuntimeError("kv cache dtype does not match kv_cache_dtype")
    num_tokens, num_heads, head_size = key.shape
    with device_of(key):
        check(lib.vltr_reshape_and_cache(
            key.data_ptr(), value.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
            slot_mapping.data_ptr(), num_tokens, num_heads, head_size, key_cache.size(3),
            key_cache.size(4), key.stride(0), value.stride(0), code, fp8, kv_scale,
            stream_of(key)))


def convert_fp8(src_cache: torch.Tensor, dst_cache: torch.Tensor) -> None:
    """csrc/cache_kernels.cu:298-339 (test helper)."""
    if not src_cache.is_cuda:
        raise RuntimeError("src must be on a GPU")
    if not dst_cache.is_cuda:
        raise RuntimeError("dst must be on a GPU")
    if src_cache.device.index != dst_cache.device.index:
        raise RuntimeError("src and dst must be on the same GPU")
    s, d = DTYPE_CODE.get(src_cache.dtype), DTYPE_CODE.get(dst_cache.dtype)
    if s is None or d is None or (s == U8) == (d == U8):
        raise RuntimeError("convert_fp8: exactly one of src/dst must be uint8")
    if not (src_cache.is_contiguous() and dst_cache.is_contiguous()):
        raise RuntimeError("convert_fp8: caches must be contiguous")
    with device_of(src_cache):
        check(lib.vltr_convert_fp8(src_cache.data_ptr(), dst_cache.data_ptr(), src_cache.numel(),
                                   s, d, stream_of(src_cache)))
