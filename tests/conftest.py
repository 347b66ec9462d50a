import glob
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tools") not in sys.path:
    sys.path.insert(1, os.path.join(ROOT, "tools"))     # bench / test drivers (hotpath.py)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load_ext(stem):
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", stem + ".*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location(stem, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cpu_has(flag):
    try:
        with open("/proc/cpuinfo") as f:
            return flag in f.read()
    except OSError:
        return False


@pytest.fixture(scope="session")
def ref_cpu():
    """The reference's own CPU kernels (csrc/cpu/*.cpp), built by oracle/Makefile into oracle/_ref."""
    import torch  # noqa: F401  (libtorch must be loaded first)
    if not _cpu_has("avx512f"):
        pytest.skip("reference CPU kernels need AVX-512")
    mod = _load_ext("ref_cpu_C" if _cpu_has("avx512_bf16") else "ref_cpu_nobf16_C")
    if mod is None:
        pytest.skip("oracle/_ref/ref_cpu_C not built (make -C oracle ref_cpu)")
    return mod


@pytest.fixture(scope="session")
def ref_cuda():
    """The reference's CUDA kernels recompiled for sm_100a (oracle/_ref/ref_cuda_C)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    mod = _load_ext("ref_cuda_C")
    if mod is None:
        pytest.skip("oracle/_ref/ref_cuda_C not built (make -C oracle ref_cuda)")
    return mod
