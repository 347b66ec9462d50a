"""GPU parity of the LTR queue sort (csrc/ltr_rank.cu through vltr_ltr_rank_host / vltr_ltr_rank)
against the reference's own expression, Python's stable `sorted` over (pri, -score) keys
(vllm/core/scheduler.py:996,998) - bit-exact permutations, including exact ties (fp16-quantised
scores as produced by the predictor), signed zeros and infinities."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def py_order(scores, pri=None, negate=True):
    sgn = -1.0 if negate else 1.0
    n = len(scores)
    if pri is None:
        return sorted(range(n), key=lambda i: sgn * scores[i])
    return sorted(range(n), key=lambda i: (pri[i], sgn * scores[i]))


@pytest.fixture(scope="module")
def ranker():
    from vllm_ltr_b200.ltr_scheduler import GpuRanker
    return GpuRanker()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 255, 256, 257, 1000, 1024, 1025, 4097, 10000, 50000])
@pytest.mark.parametrize("with_pri", [False, True])
@pytest.mark.parametrize("quant", ["fp16", "coarse", "f64"])
def test_rank_matches_python_sorted(ranker, n, with_pri, quant):
    rng = np.random.default_rng(n + 7)
    s = rng.normal(0, 1, n)
    if quant == "fp16":
        s = s.astype(np.float16).astype(np.float64)       # what logits[:,0].tolist() yields
    elif quant == "coarse":
        s = np.round(s * 2) / 2                            # few distinct values: long tie runs
    pri = rng.choice([-1, 0, 0, 0, 3], n).astype(np.int32) if with_pri else None
    got = ranker.order(s.tolist(), None if pri is None else pri.tolist())
    assert got.tolist() == py_order(s.tolist(), None if pri is None else pri.tolist())


def test_rank_special_values(ranker):
    s = [0.0, -0.0, 1.5, -1.5, float("inf"), -float("inf"), 0.0, -0.0, 1e-320, -1e-320, 5e-324,
         1.5, float("inf")]
    for negate in (True, False):
        assert ranker.order(s, None, negate).tolist() == py_order(s, None, negate)
    pri = [0, 0, -1, 0, 0, -1, 0, -1, 0, 0, 0, -1, -1]
    assert ranker.order(s, pri).tolist() == py_order(s, pri)


def test_rank_is_a_permutation_at_scale(ranker):
    rng = np.random.default_rng(0)
    n = 200_000
    s = rng.normal(0, 1, n).astype(np.float16).astype(np.float64)
    got = ranker.order(s)
    assert np.array_equal(np.sort(got), np.arange(n))
    key = -s[got]
    assert (np.diff(key) >= 0).all()
    same = np.diff(key) == 0
    assert (np.diff(got)[same] > 0).all()          # ties keep input order


def test_rank_device_entry_point():
    """vltr_ltr_rank on device arrays (the asynchronous entry of include/vllm_ltr_b200.h)."""
    from vllm_ltr_b200._lib import check, lib
    n = 5000
    g = torch.Generator(device="cuda:0").manual_seed(1)
    s = torch.randn(n, generator=g, device="cuda:0", dtype=torch.float64).half().double()
    pri = torch.randint(-1, 1, (n,), generator=g, device="cuda:0", dtype=torch.int32)
    order = torch.empty(n, dtype=torch.int32, device="cuda:0")
    wsb = lib.vltr_ltr_rank_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda:0")
    stream = torch._C._cuda_getCurrentRawStream(0)
    check(lib.vltr_ltr_rank(s.data_ptr(), pri.data_ptr(), n, 1, order.data_ptr(), ws.data_ptr(),
                            ctypes.c_size_t(wsb), stream))
    torch.cuda.synchronize()
    assert order.cpu().tolist() == py_order(s.cpu().tolist(), pri.cpu().tolist())
    with pytest.raises(RuntimeError, match="workspace too small"):
        check(lib.vltr_ltr_rank(s.data_ptr(), pri.data_ptr(), n, 1, order.data_ptr(),
                                ws.data_ptr(), ctypes.c_size_t(16), stream))
