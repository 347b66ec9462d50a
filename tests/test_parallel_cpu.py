"""world_size-2 `gloo` test of the tensor-parallel exchange wrappers (vllm_ltr_b200/parallel.py
and the vllm._C.custom_ar surface): integer-valued tensors so the sum must be exact, like the
reference's tests/distributed/test_custom_all_reduce.py:30-51."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vllm_ltr_b200 import parallel
        from vllm_ltr_b200._C import custom_ar
        parallel.init_tensor_parallel()
        assert parallel.get_tensor_model_parallel_world_size() == world
        g = torch.Generator().manual_seed(rank)
        x = torch.randint(-8, 8, (256, 512), generator=g).to(torch.bfloat16)
        want = sum(torch.randint(-8, 8, (256, 512), generator=torch.Generator().manual_seed(r))
                   for r in range(world)).to(torch.bfloat16)
        y = parallel.tensor_model_parallel_all_reduce(x.clone())
        assert torch.equal(y, want)
        ptr = custom_ar.init_custom_ar(torch.zeros(custom_ar.meta_size() + 1024, dtype=torch.uint8),
                                       torch.empty(1024, dtype=torch.uint8), ["h"] * world,
                                       [0] * world, rank, True)
        out = torch.empty_like(x)
        custom_ar.all_reduce_unreg(ptr, x, torch.empty(x.numel() * 2, dtype=torch.uint8), out)
        assert torch.equal(out, want) and not torch.equal(x, want)      # out of place
        out2 = torch.empty_like(x)
        custom_ar.all_reduce_reg(ptr, x, out2)
        assert torch.equal(out2, want)
        try:
            custom_ar.all_reduce_unreg(ptr, x, torch.empty(16, dtype=torch.uint8), out)
            raise AssertionError("too-small registered buffer must be rejected")
        except RuntimeError:
            pass
        assert custom_ar.should_custom_ar(x, 8 << 20, world, True) is False   # CPU tensor
        custom_ar.dispose(ptr)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_all_reduce_world2_gloo():
    # spawned children re-import this module by name with the parent's sys.path: the repo root
    # must come before /root/reference (which other tests put first and which has its own
    # `tests` package)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_init_custom_ar_argument_checks():
    from vllm_ltr_b200._C import custom_ar
    with pytest.raises(ValueError, match="Odd num gpus"):
        custom_ar.init_custom_ar(None, None, ["a"] * 3, [0] * 3, 0, True)
    with pytest.raises(ValueError, match="world size > 8"):
        custom_ar.init_custom_ar(None, None, ["a"] * 10, [0] * 10, 0, True)
    with pytest.raises(ValueError, match="invalid rank"):
        custom_ar.init_custom_ar(None, None, ["a"] * 2, [0] * 2, 5, True)
