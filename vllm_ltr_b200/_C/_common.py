"""Helpers shared by the op wrappers: dtype codes, stream lookup, argument checks."""
import torch

from .._lib import check, lib  # noqa: F401

F32, F16, BF16, U8 = 0, 1, 2, 3
DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16, torch.uint8: U8}

_raw_stream = torch._C._cuda_getCurrentRawStream


def stream_of(t: torch.Tensor) -> int:
    """Raw cudaStream_t of torch's current stream on the tensor's device (the reference launches on
    at::cuda::getCurrentCUDAStream() under an OptionalCUDAGuard of the first tensor)."""
    return _raw_stream(t.device.index)


class device_of:
    """OptionalCUDAGuard(device_of(tensor)) — only switches when the tensor is on another GPU."""
    __slots__ = ("idx", "prev")

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def float_code(t: torch.Tensor, what: str) -> int:
    code = DTYPE_CODE.get(t.dtype)
    if code is None or code == U8:
        raise RuntimeError(f"Unsupported data type: {t.dtype} ({what})")
    return code


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} must be a CUDA tensor: vllm_ltr_b200 has no CPU path (got {t.device})")
