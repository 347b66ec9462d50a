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


# ---------------------------------------------------------------------------------------------
# fused exchange + norm over NVLink peer memory (csrc/tp_allreduce.cu)
# ---------------------------------------------------------------------------------------------
class FusedTPComm:
    """Symmetric activation buffers of one tensor-parallel rank and the fused
    ``all_reduce + residual add + RMSNorm`` that replaces ``tensor_model_parallel_all_reduce``
    (communication_op.py:13-42) followed by ``fused_add_rms_norm`` (llama.py:236-238).

    ``x[i]`` are the buffers the row-parallel GEMMs (o_proj, down_proj) write their partial sums
    into, ``out`` receives the normalised rows and ``resid`` is the running residual - all
    allocated from ``torch.distributed._symmetric_memory`` so every rank holds every rank's
    pointers (plumbing only; the data path is the kernel in libvllm_ltr_b200).  When the
    symmetric-memory handles carry multicast mappings (NVLS), the kernel sums rows inside the
    NVSwitch (``multimem.ld_reduce``) and publishes them with ``multimem.st``
    (``multimem=False`` / ``VLTR_TP_MULTIMEM=0`` keeps peer loads and stores).

    ``replicate_residual=False`` (default): the running residual of a row lives on the rank that
    owns the row (``row % world == rank``) - the only rank that ever reads it again - so
    ``resid`` is valid on a rank only for the rows it owns and an exchange publishes half the
    bytes; ``True`` keeps the reference's replicated residual on every rank."""

    def __init__(self, max_tokens: int, hidden: int, dtype: torch.dtype, device,
                 group: Optional[dist.ProcessGroup] = None, num_inputs: int = 2,
                 device_epoch: bool = True, multimem: Optional[bool] = None,
                 replicate_residual: bool = False):
        import ctypes
        import os

        import torch.distributed._symmetric_memory as symm
        from ._lib import declare
        self._ct = ctypes
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world < 2 or self.world > 8:
            raise RuntimeError(f"FusedTPComm supports 2..8 ranks, got {self.world}")
        self.hidden, self.dtype, self.max_tokens = hidden, dtype, max_tokens
        name = group.group_name
        mk = lambda: symm.empty(max_tokens, hidden, dtype=dtype, device=device)
        self.x = [mk() for _ in range(num_inputs)]
        self.out, self.resid = mk(), mk()
        # the rendezvous flags live in a buffer of our own (zeroed, 16 KiB; the kernels use u32
        # indices 1024..1040), not in torch's signal pads: no assumption about their size or
        # about which part of them torch's own barriers use
        self.flags = symm.empty(4096, dtype=torch.int32, device=device)
        self.flags.zero_()
        hx = [symm.rendezvous(t, name) for t in self.x]
        ho, hr = symm.rendezvous(self.out, name), symm.rendezvous(self.resid, name)
        hf = symm.rendezvous(self.flags, name)
        self._handles = hx + [ho, hr, hf]
        arr = lambda ptrs: (ctypes.c_void_p * self.world)(*[int(p) for p in ptrs])
        self._x_ptrs = [arr(h.buffer_ptrs) for h in hx]
        self._out_ptrs, self._resid_ptrs = arr(ho.buffer_ptrs), arr(hr.buffer_ptrs)
        self._pads = arr(hf.buffer_ptrs)
        # NVLS: multicast mappings of the same buffers when the fabric offers them (0 otherwise)
        mcp = lambda h: int(getattr(h, "multicast_ptr", 0) or 0)
        self._x_mc = [mcp(h) for h in hx]
        self._out_mc, self._resid_mc = mcp(ho), mcp(hr)
        want_mc = multimem if multimem is not None else os.environ.get("VLTR_TP_MULTIMEM", "1") != "0"
        self.multimem = bool(want_mc and all(self._x_mc) and self._out_mc and self._resid_mc)

        class _Comm(ctypes.Structure):
            _fields_ = [("rank", ctypes.c_int), ("world", ctypes.c_int),
                        ("epoch", ctypes.c_uint32), ("x_ptrs", ctypes.c_void_p), ("out_ptrs", ctypes.c_void_p),
                        ("resid_ptrs", ctypes.c_void_p), ("signal_pads", ctypes.c_void_p),
                        ("x_mc", ctypes.c_void_p), ("out_mc", ctypes.c_void_p),
                        ("resid_mc", ctypes.c_void_p), ("resid_local", ctypes.c_int)]

        addr = lambda a: ctypes.cast(a, ctypes.c_void_p)
        # device_epoch: the kernels count the exchanges themselves (in the signal pad), so the
        # launch arguments never change and a step can be replayed from a CUDA graph; otherwise
        # the host passes 2 * call_index + 1
        self.device_epoch = device_epoch
        self.replicate_residual = replicate_residual
        self._calls = 0       # exchanges issued on these buffers (same on every rank)
        mc = self.multimem
        self._comms = [_Comm(self.rank, self.world, 1, addr(xp), addr(self._out_ptrs),
                             addr(self._resid_ptrs), addr(self._pads),
                             self._x_mc[i] if mc else None, self._out_mc if mc else None,
                             self._resid_mc if mc else None, 0 if replicate_residual else 1)
                       for i, xp in enumerate(self._x_ptrs)]
        _P = ctypes.c_void_p
        self._norm = declare("vltr_tp_allreduce_add_rms_norm", ctypes.c_int,
                             [_P, _P, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P])
        self._plain = declare("vltr_tp_allreduce", ctypes.c_int,
                              [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P])
        self._code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dtype]
        dist.barrier(group)

    def _stream(self):
        return torch._C._cuda_getCurrentRawStream(self.out.device.index)

    def allreduce_add_rms_norm(self, which: int, weight: torch.Tensor, eps: float,
                               num_tokens: int) -> torch.Tensor:
        """resid = (T)(sum over ranks of x[which]) + resid; returns out = rms_norm(resid) * weight
        (rows [:num_tokens]); identical on every rank."""
        from ._lib import check
        self._comms[which].epoch = 0 if self.device_epoch else (2 * self._calls + 1) & 0xFFFFFFFF
        self._calls += 1
        check(self._norm(self._ct.byref(self._comms[which]), weight.data_ptr(), eps, num_tokens,
                         self.hidden, self._code, self._stream()))
        return self.out[:num_tokens]

    def allreduce(self, which: int, num_tokens: int) -> torch.Tensor:
        """out = sum over ranks of x[which] (rows [:num_tokens])."""
        from ._lib import check
        self._comms[which].epoch = 0 if self.device_epoch else (2 * self._calls + 1) & 0xFFFFFFFF
        self._calls += 1
        check(self._plain(self._ct.byref(self._comms[which]), num_tokens, self.hidden, self._code,
                          self._stream()))
        return self.out[:num_tokens]
