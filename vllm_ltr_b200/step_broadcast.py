"""Driver -> workers broadcast of one step's inputs as ONE buffer (tensor-parallel engines).

Drop-in for ``vllm.distributed.broadcast_tensor_dict`` (vllm/distributed/communication_op.py:
143-211), called by ``ModelRunner.prepare_input_tensors`` (worker/model_runner.py:760-770, twice
for a mixed prefill + decode step) and ``Worker.execute_model`` (worker/worker.py:234-236).  The
reference pickles the key / dtype / shape list through ``broadcast_object_list`` (two collectives
with host synchronisation: a size, then the bytes) and then issues one broadcast PER TENSOR -
about a dozen latency-bound collectives per step at 8 GPUs (its own comment: "costs 300us
overhead ... We can potentially reduce the overhead by coelescing tensors", model_runner.py:
764-765).

Here the dictionary travels as one byte buffer:

    [ int64 header_bytes | int64 total_bytes | pickled entries | 16-byte aligned tensor payloads ]

* one broadcast of the first ``HEAD_BYTES`` (header + the payloads of almost every decode step);
* a second broadcast of the remainder only when the step does not fit (its size is in the head);
* receivers get VIEWS into the staging buffer (no per-tensor allocation or copy); staging
  buffers rotate (``RING``) so the views of one step stay valid while the next is received.

Same signature, same return value (the driver gets its own dictionary back, workers a new one
with the same keys in the same order), same constraint that tensors live on the step's device.
"""
import pickle
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.distributed as dist

HEAD_BYTES = 256 << 10      # first broadcast: header + payload prefix
_PREFIX = 16                # two int64: header length, total length
RING = 4                    # staging buffers in rotation (a mixed step broadcasts twice)

_state: Dict[Tuple[int, str], "_Staging"] = {}


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class _Staging:
    def __init__(self, device: torch.device):
        self.device = device
        self.bufs: List[torch.Tensor] = [torch.empty(HEAD_BYTES, dtype=torch.uint8, device=device)
                                         for _ in range(RING)]
        # host side of the header, one per staging buffer: the host -> device copy of step k is
        # asynchronous, the engine synchronises with the device once per step (sampler output)
        self.hosts: List[torch.Tensor] = [self._host(16 << 10) for _ in range(RING)]
        self.cursor = 0

    def _host(self, nbytes: int) -> torch.Tensor:
        return torch.empty(_align(nbytes, 1 << 12), dtype=torch.uint8,
                           pin_memory=self.device.type == "cuda")

    def next(self, nbytes: int) -> torch.Tensor:
        i = self.cursor
        self.cursor = (i + 1) % RING
        self.host_index = i
        if self.bufs[i].numel() < nbytes:
            self.bufs[i] = torch.empty(_align(nbytes, 1 << 16), dtype=torch.uint8,
                                       device=self.device)
        return self.bufs[i]

    def host(self, nbytes: int) -> torch.Tensor:
        """Host staging paired with the buffer `next` handed out last."""
        i = self.host_index
        if self.hosts[i].numel() < nbytes:
            self.hosts[i] = self._host(nbytes)
        return self.hosts[i]


def _staging(group, device: torch.device) -> _Staging:
    key = (id(group), str(device))
    st = _state.get(key)
    if st is None:
        st = _state[key] = _Staging(device)
    return st


def _default_device(group) -> torch.device:
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_tensor_dict(
    tensor_dict: Optional[Dict[Any, Union[torch.Tensor, Any]]] = None,
    src: int = 0,
    group=None,
) -> Optional[Dict[Any, Union[torch.Tensor, Any]]]:
    group = group or dist.group.WORLD
    ranks = dist.get_process_group_ranks(group)
    assert src in ranks, f"Invalid src rank ({src})"
    if dist.get_world_size(group=group) == 1:
        return tensor_dict
    device = _default_device(group)
    st = _staging(group, device)

    if dist.get_rank() == src:
        assert isinstance(tensor_dict, dict), f"Expecting a dictionary, got {type(tensor_dict)}"
        entries: List[Tuple[Any, Any]] = []
        tensors: List[Tuple[torch.Tensor, int, int]] = []
        off = 0                                    # payload offsets are relative to the payload start
        for key, value in tensor_dict.items():
            if isinstance(value, torch.Tensor):
                assert value.device.type == device.type, (
                    f"Tensor {key} is on {value.device}; the step's tensors must be on {device}")
                nbytes = value.numel() * value.element_size()
                entries.append((key, ("__tensor__", value.dtype, tuple(value.shape), off, nbytes)))
                tensors.append((value, off, nbytes))
                off = _align(off + nbytes)
            else:
                entries.append((key, value))
        header = pickle.dumps(entries, protocol=pickle.HIGHEST_PROTOCOL)
        base = _align(_PREFIX + len(header))
        total = base + off
        buf = st.next(max(total, HEAD_BYTES))
        # header: host staging -> device (one copy), payloads: device -> device
        host = st.host(base)
        host[:_PREFIX].view(torch.int64).copy_(torch.tensor([len(header), total]))
        host[_PREFIX:_PREFIX + len(header)].copy_(
            torch.frombuffer(bytearray(header), dtype=torch.uint8))
        buf[:base].copy_(host[:base], non_blocking=True)
        for t, o, nbytes in tensors:
            if nbytes:
                buf[base + o:base + o + nbytes].copy_(
                    t.contiguous().view(-1).view(torch.uint8), non_blocking=True)
        dist.broadcast(buf[:HEAD_BYTES], src=src, group=group)
        if total > HEAD_BYTES:
            dist.broadcast(buf[HEAD_BYTES:total], src=src, group=group)
        return tensor_dict

    buf = st.next(HEAD_BYTES)
    dist.broadcast(buf[:HEAD_BYTES], src=src, group=group)
    # the header is needed on the host: one device -> host copy of the head's first bytes
    probe = 16 << 10
    host = st.host(probe)
    host[:probe].copy_(buf[:probe])
    hlen, total = (int(x) for x in host[:_PREFIX].view(torch.int64).tolist())
    base = _align(_PREFIX + hlen)
    if total > HEAD_BYTES:
        if buf.numel() < total:                    # keep the head, receive the rest behind it
            big = torch.empty(_align(total, 1 << 16), dtype=torch.uint8, device=device)
            big[:HEAD_BYTES].copy_(buf[:HEAD_BYTES])
            st.bufs[st.host_index] = big
            buf = big
        dist.broadcast(buf[HEAD_BYTES:total], src=src, group=group)
    if base > probe:                               # an unusually long header: fetch all of it
        host = st.host(base)
        host[:base].copy_(buf[:base])
    entries = pickle.loads(host[_PREFIX:_PREFIX + hlen].numpy().tobytes())
    out: Dict[Any, Any] = {}
    for key, value in entries:
        if isinstance(value, tuple) and len(value) == 5 and value[0] == "__tensor__":
            _, dtype, shape, o, nbytes = value
            if nbytes:
                out[key] = buf[base + o:base + o + nbytes].view(dtype).view(shape)
            else:
                out[key] = torch.empty(shape, dtype=dtype, device=device)
        else:
            out[key] = value
    return out


def install() -> List[str]:
    """Re-bind the reference's call sites (they import the function by name)."""
    import importlib
    done = []
    for modname in ("vllm.distributed.communication_op", "vllm.distributed",
                    "vllm.worker.model_runner", "vllm.worker.worker"):
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        if hasattr(mod, "broadcast_tensor_dict"):
            setattr(mod, "broadcast_tensor_dict", broadcast_tensor_dict)
            done.append(modname)
    return done
