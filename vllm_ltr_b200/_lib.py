"""ctypes loader for libvllm_ltr_b200.so (the C-ABI declared in include/vllm_ltr_b200.h).

The CUDA library IS the product: there is no Python/torch fallback.  If the shared object is
missing, or was built for anything but sm_100a, importing this module raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# VLTR_LIB: development builds of the same library (e.g. the -DVLTR_ATTN_TRACE=1 variant that
# tools/attn_tune.py --trace needs); the default is the in-tree product build
LIB_PATH = os.environ.get("VLTR_LIB") or os.path.join(_HERE, "libvllm_ltr_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C vllm_ltr_b200/csrc` (nvcc, sm_100a). There is no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

_P = c_void_p
_SIGS = {
    "vltr_last_error": (c_char_p, []),
    "vltr_build_info": (c_char_p, []),
    "vltr_launch_count": (c_uint64, []),
    "vltr_tune": (c_int, [c_char_p, ctypes.c_longlong]),
    "vltr_paged_attention_v1": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P,
                                        c_int, c_int, c_int, _P, c_int64, c_int64, c_int64, c_int,
                                        c_int, c_float, _P]),
    "vltr_paged_attention_v2": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int,
                                        c_float, _P, _P, c_int, c_int, c_int, _P, c_int64, c_int64,
                                        c_int64, c_int, c_int, c_float, _P]),
    "vltr_reshape_and_cache": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                       c_int64, c_int64, c_int, c_int, c_float, _P]),
    "vltr_copy_blocks": (c_int, [_P, _P, c_int, _P, c_int, c_int64, _P]),
    "vltr_swap_blocks": (c_int, [_P, _P, c_int, _P, c_int, c_int64, _P]),
    "vltr_swap_blocks_layers": (c_int, [_P, _P, c_int, c_int, _P, c_int, c_int64, _P]),
    "vltr_convert_fp8": (c_int, [_P, _P, c_int64, c_int, c_int, _P]),
    "vltr_rms_norm": (c_int, [_P, _P, _P, c_float, c_int, c_int, c_int, _P]),
    "vltr_fused_add_rms_norm": (c_int, [_P, _P, _P, c_float, c_int, c_int, c_int, _P]),
    "vltr_rotary_embedding": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int64,
                                      c_int64, _P, _P, c_int, c_int, _P]),
    "vltr_activation": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P]),
    "vltr_ltr_rank_workspace_bytes": (ctypes.c_size_t, [c_int]),
    "vltr_ltr_rank": (c_int, [_P, _P, c_int, c_int, _P, _P, ctypes.c_size_t, _P]),
    "vltr_ltr_rank_host": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "vltr_get_device_attribute": (c_int, [c_int, c_int]),
    "vltr_get_max_shared_memory_per_block_device_attribute": (c_int, [c_int]),
}

# symbols added by later translation units are registered here by the modules that own them
def declare(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


for _name, (_res, _args) in _SIGS.items():
    declare(_name, _res, _args)

_info = lib.vltr_build_info().decode()
if not _info.startswith("sm_100a"):
    raise ImportError(f"libvllm_ltr_b200.so was not built for sm_100a: {_info}")


def last_error() -> str:
    return lib.vltr_last_error().decode(errors="replace")


def check(rc: int) -> None:
    """Turn a VLTR_ERR_* code into the RuntimeError the reference's TORCH_CHECK would raise."""
    if rc != 0:
        raise RuntimeError(last_error())


def launch_count() -> int:
    return int(lib.vltr_launch_count())


def tune(key: str, value: int) -> None:
    """Development knobs of the library (include/vllm_ltr_b200.h: vltr_tune)."""
    check(lib.vltr_tune(key.encode(), int(value)))
