"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/vllm_ltr_b200.h declares, contains sm_100a code only, and the vllm._C mirror exposes the
reference's pybind surface (csrc/pybind.cpp:6-125).  No compute calls (there is no GPU here)."""
import ctypes
import inspect
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vllm_ltr_b200.h")
LIB = os.path.join(ROOT, "vllm_ltr_b200", "libvllm_ltr_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vltr_\w+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("vltr_paged_attention_v1", "vltr_paged_attention_v2", "vltr_reshape_and_cache",
                 "vltr_copy_blocks", "vltr_swap_blocks", "vltr_convert_fp8", "vltr_rms_norm",
                 "vltr_fused_add_rms_norm", "vltr_rotary_embedding", "vltr_activation"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_library_is_sm100a_only():
    lib = ctypes.CDLL(LIB)
    lib.vltr_build_info.restype = ctypes.c_char_p
    assert lib.vltr_build_info().decode().startswith("sm_100a")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_torch_types_in_the_abi():
    src = open(HEADER).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S).lower()
    assert 'extern "C"' in src


def test_header_compiles_as_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "vllm_ltr_b200.h"\nint main(void){return VLTR_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), "-c",
                    str(c), "-o", str(tmp_path / "t.o")], check=True)


def test_mirror_surface_matches_reference_pybind():
    """names + arity of csrc/pybind.cpp (ops.h / cache.h / cuda_utils.h signatures)."""
    from vllm_ltr_b200 import _C
    want = {
        _C.ops: {"paged_attention_v1": 13, "paged_attention_v2": 16, "silu_and_mul": 2,
                 "gelu_and_mul": 2, "gelu_tanh_and_mul": 2, "gelu_new": 2, "gelu_fast": 2,
                 "rms_norm": 4, "fused_add_rms_norm": 4, "rotary_embedding": 6,
                 "batched_rotary_embedding": 8},
        _C.cache_ops: {"swap_blocks": 3, "copy_blocks": 3, "reshape_and_cache": 7,
                       "convert_fp8": 2},
        _C.cuda_utils: {"get_device_attribute": 2,
                        "get_max_shared_memory_per_block_device_attribute": 1},
    }
    for mod, fns in want.items():
        for name, arity in fns.items():
            fn = getattr(mod, name)
            assert len(inspect.signature(fn).parameters) == arity, name


def test_install_registers_vllm_C():
    import sys
    import vllm_ltr_b200
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith("vllm._C")}
    try:
        vllm_ltr_b200.install()
        from vllm_ltr_b200 import _C
        assert sys.modules["vllm._C"] is _C
        assert sys.modules["vllm._C.ops"] is _C.ops
    finally:
        for k in [k for k in sys.modules if k.startswith("vllm._C")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: a CPU tensor raises instead of being computed somewhere else."""
    import torch
    from vllm_ltr_b200 import _C
    with pytest.raises(RuntimeError, match="no CPU path"):
        _C.ops.silu_and_mul(torch.empty(2, 4), torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        _C.ops.rms_norm(torch.empty(2, 8), torch.zeros(2, 8), torch.ones(8), 1e-5)
