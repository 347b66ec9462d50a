"""Tensor-parallel exchange step of the decode hot path (config 5: 70B at TP=8).

The path shards naturally by attention head / MLP column (vllm/model_executor/models/llama.py:
96-111, layers/linear.py); the only exchange is the all-reduce(sum) of the [tokens, hidden]
row-parallel GEMM outputs - two per decoder layer (o_proj, down_proj; linear.py:577) plus the
embedding.  The reference sends those through ``tensor_model_parallel_all_reduce``
(vllm/distributed/communication_op.py:13-42): its custom CUDA-IPC kernel when the message is
small and the GPUs are fully NVLinked, else NCCL.

Here: one process per GPU, ``torch.distributed`` (backend ``nccl``) over NVLink 5 / NVSwitch.
On an NVSwitch domain every peer is one hop away at full bandwidth and NCCL can reduce inside
the switch (NVLS), so the one-stage / two-stage split of the reference kernel has no
counterpart; the ``vllm._C.custom_ar`` surface (``vllm_ltr_b200._C.custom_ar``) is kept and
routed here so the reference's ``CustomAllreduce`` wrapper keeps working unchanged.
"""
from typing import Optional

import torch
import torch.distributed as dist

_GROUP: Optional[dist.ProcessGroup] = None


def init_tensor_parallel(group: Optional[dist.ProcessGroup] = None) -> None:
    """Use `group` (default: the world group) for the tensor-parallel all-reduce."""
    global _GROUP
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _GROUP = group


def get_tensor_model_parallel_world_size() -> int:
    return dist.get_world_size(_GROUP) if dist.is_initialized() else 1


def get_tensor_model_parallel_rank() -> int:
    return dist.get_rank(_GROUP) if dist.is_initialized() else 0


def tensor_model_parallel_all_reduce(input_: torch.Tensor) -> torch.Tensor:
    """communication_op.py:13-42: sums `input_` over the tensor-parallel group.  In place;
    returns the tensor (callers must use the return value, like the reference's)."""
    if get_tensor_model_parallel_world_size() == 1:
        return input_
    dist.all_reduce(input_, op=dist.ReduceOp.SUM, group=_GROUP)
    return input_


def all_reduce_into(inp: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """Out-of-place form used by the custom_ar surface (csrc/custom_all_reduce.cu:101-122)."""
    if inp.dtype != out.dtype or inp.numel() != out.numel():
        raise RuntimeError("all_reduce: inp and out must match in dtype and size")
    if out.data_ptr() != inp.data_ptr():
        out.copy_(inp)
    return tensor_model_parallel_all_reduce(out)
