"""vllm_ltr_b200 — B200 (sm_100a) drop-in for the decode-time hot path of hao-ai-lab/vllm-ltr.

Public surface
--------------
vllm_ltr_b200._C           mirror of the reference's ``vllm._C`` pybind module
                           (``ops``, ``cache_ops``, ``cuda_utils``, ``custom_ar``) over the C-ABI
                           library ``libvllm_ltr_b200.so`` (``include/vllm_ltr_b200.h``)
vllm_ltr_b200.install()    registers that mirror as ``sys.modules["vllm._C"]`` so the unmodified
                           reference engine (``vllm/_custom_ops.py:5-9``) picks it up
"""
import sys

__version__ = "0.1.0"


def install() -> None:
    """Make ``from vllm._C import ops, cache_ops, cuda_utils, custom_ar`` resolve to this package.

    Call before ``import vllm`` (or at least before ``vllm._custom_ops`` is imported).
    """
    from . import _C
    sys.modules["vllm._C"] = _C
    for sub in ("ops", "cache_ops", "cuda_utils", "custom_ar"):
        sys.modules[f"vllm._C.{sub}"] = getattr(_C, sub)
    vllm = sys.modules.get("vllm")
    if vllm is not None:
        setattr(vllm, "_C", _C)
