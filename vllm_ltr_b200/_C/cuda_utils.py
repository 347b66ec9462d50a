"""``vllm._C.cuda_utils`` (csrc/pybind.cpp:96-106, csrc/cuda_utils_kernels.cu)."""
from .._lib import lib


def get_device_attribute(attribute: int, device_id: int) -> int:
    return int(lib.vltr_get_device_attribute(attribute, device_id))


def get_max_shared_memory_per_block_device_attribute(device_id: int) -> int:
    return int(lib.vltr_get_max_shared_memory_per_block_device_attribute(device_id))
