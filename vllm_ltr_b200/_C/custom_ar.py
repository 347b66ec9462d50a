"""``vllm._C.custom_ar`` - tensor-parallel all-reduce surface (csrc/pybind.cpp:108-123,
csrc/ops.h:143-160, implementation csrc/custom_all_reduce.cu/.cuh of the reference).

The reference's Python wrapper (vllm/distributed/device_communicators/custom_all_reduce.py:
163-260) drives this surface: it allocates a metadata + staging buffer, exchanges CUDA-IPC handles,
calls ``init_custom_ar`` and then ``all_reduce_unreg`` / ``all_reduce_reg`` per message.  The same
calls work here; the exchange itself goes over NCCL on the tensor-parallel group
(vllm_ltr_b200/parallel.py) - on B200 every GPU pair is one NVSwitch hop apart, so there are no
IPC pointers to register and no 8 MiB / world-size limits:

* ``init_custom_ar`` records (rank, world size) and returns an opaque handle; the IPC handles and
  offsets it is given are accepted and ignored;
* ``should_custom_ar`` keeps the reference's shape checks (16-byte multiple, contiguity,
  csrc/custom_all_reduce.cu:43-70) but not its size caps;
* ``register_buffer`` / ``register_graph_buffers`` / ``get_graph_buffer_ipc_meta`` are no-ops
  with the reference's return types (NCCL manages its own registration).
"""
from typing import Dict, List, Tuple

import torch

from .. import parallel

_HANDLES: Dict[int, Tuple[int, int]] = {}
_NEXT = [1]


def meta_size() -> int:
    """sizeof(vllm::Signal) in the reference (csrc/custom_all_reduce.cuh:26-36): start/end
    flag arrays of 8 ranks x 2 - kept so the wrapper's buffer arithmetic is unchanged."""
    return 256


def init_custom_ar(meta: torch.Tensor, rank_data: torch.Tensor, handles: List[str],
                   offsets: List[int], rank: int, full_nvlink: bool) -> int:
    world_size = len(offsets)
    if world_size > 8:
        raise ValueError("world size > 8 is not supported")
    if world_size % 2 != 0:
        raise ValueError("Odd num gpus is not supported for now")
    if world_size != len(handles):
        raise ValueError("handles length should equal to offsets length")
    if rank < 0 or rank >= world_size:
        raise ValueError("invalid rank passed in")
    if parallel.get_tensor_model_parallel_world_size() != world_size:
        raise RuntimeError("init_custom_ar: tensor-parallel group size does not match the "
                           f"number of handles ({world_size})")
    ptr = _NEXT[0]
    _NEXT[0] += 1
    _HANDLES[ptr] = (rank, world_size)
    return ptr


def should_custom_ar(inp: torch.Tensor, max_size: int, world_size: int,
                     full_nvlink: bool) -> bool:
    inp_size = inp.numel() * inp.element_size()
    if inp_size % 16 != 0:
        return False
    if not inp.is_contiguous():
        return False
    return inp.is_cuda


def _check(ptr: int) -> None:
    if ptr not in _HANDLES:
        raise RuntimeError("custom_ar: invalid or disposed handle")


def all_reduce_reg(ptr: int, inp: torch.Tensor, out: torch.Tensor) -> None:
    _check(ptr)
    parallel.all_reduce_into(inp, out)


def all_reduce_unreg(ptr: int, inp: torch.Tensor, reg_buffer: torch.Tensor,
                     out: torch.Tensor) -> None:
    _check(ptr)
    if inp.numel() * inp.element_size() > reg_buffer.numel() * reg_buffer.element_size():
        raise RuntimeError("registered buffer is too small to contain the input")
    parallel.all_reduce_into(inp, out)     # no staging copy: NCCL reads the input in place


def dispose(ptr: int) -> None:
    _HANDLES.pop(ptr, None)


def register_buffer(ptr: int, t: torch.Tensor, handles: List[str], offsets: List[int]) -> None:
    _check(ptr)


def get_graph_buffer_ipc_meta(ptr: int) -> Tuple[List[int], List[int]]:
    _check(ptr)
    return [], []


def register_graph_buffers(ptr: int, handles: List[str], offsets: List[List[int]]) -> None:
    _check(ptr)
