"""Mirror of the reference's ``vllm._C`` extension module (csrc/pybind.cpp:6-125).

Submodules and function names, argument order and in-place semantics are the reference's; the
work is done by libvllm_ltr_b200.so through the C ABI of include/vllm_ltr_b200.h.
"""
from . import cache_ops, cuda_utils, custom_ar, ops  # noqa: F401

__all__ = ["ops", "cache_ops", "cuda_utils", "custom_ar"]
