"""world_size-2 `gloo` test of the coalesced step-input broadcast (vllm_ltr_b200/step_broadcast.py,
drop-in for vllm/distributed/communication_op.py:143-211): workers must receive the driver's
dictionary - same keys in the same order, equal tensors (dtype, shape, values), equal Python
values - through at most two collectives per call, for a decode step, a step too large for the
head buffer, an empty tensor and consecutive calls whose results stay valid (buffer ring)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dicts():
    g = torch.Generator().manual_seed(0)
    B = 256
    decode = {
        "input_tokens": torch.randint(0, 32000, (B,), generator=g),
        "input_positions": torch.randint(0, 4096, (B,), generator=g),
        "selected_token_indices": torch.arange(B),
        "lora_requests": set(), "lora_mapping": None, "multi_modal_input": None,
        "num_prefill_tokens": 0, "num_decode_tokens": B,
        "slot_mapping": torch.randint(0, 1 << 20, (B,), generator=g),
        "num_prefills": 0, "batch_type": ("enum", 2), "need_score": False,
        "context_lens": torch.randint(1, 2048, (B,), generator=g).to(torch.int32),
        "block_tables": torch.randint(0, 18000, (B, 140), generator=g).to(torch.int32),
        "kv_cache_dtype": "auto", "use_cuda_graph": False, "max_context_len": 2047,
    }
    big = {"block_tables": torch.randint(0, 18000, (256, 512), generator=g).to(torch.int32),
           "x": torch.randn(3, 5, 7, generator=g).to(torch.bfloat16), "n": 7,
           "empty": torch.empty(0, 4, dtype=torch.float16),
           "flags": torch.tensor([True, False, True]), "scalar": torch.tensor(3.5)}
    small = {"a": torch.arange(5, dtype=torch.int16), "b": "text", 3: [1, 2, {"k": None}]}
    return [decode, big, small, decode]


def _same(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        if isinstance(a[k], torch.Tensor):
            assert isinstance(b[k], torch.Tensor) and a[k].dtype == b[k].dtype, k
            assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vllm_ltr_b200 import step_broadcast as sb
        calls = []
        real = dist.broadcast

        def counting(t, src, group=None, async_op=False):
            calls.append(t.numel())
            return real(t, src=src, group=group, async_op=async_op)

        sb.dist.broadcast = counting
        want = _dicts()
        got = []
        for d in want:
            n0 = len(calls)
            out = sb.broadcast_tensor_dict(d if rank == 0 else None, src=0)
            if rank == 0:
                assert out is d
            got.append(out)
            assert len(calls) - n0 <= 2, calls[n0:]
        # the decode step fits the head buffer: exactly one collective
        assert calls[0] == sb.HEAD_BYTES and calls[1] != sb.HEAD_BYTES or len(calls) >= 5
        # every result is still intact after the later calls (ring of staging buffers)
        for a, b in zip(want, got):
            _same(a, b)
        q.put((rank, "ok", len(calls)))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), 0))
    finally:
        dist.destroy_process_group()


def test_step_broadcast_world2_gloo():
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
    assert sorted(r[:2] for r in res) == [(0, "ok"), (1, "ok")], res
    # decode (1) + big (2) + small (1) + decode (1) collectives on every rank
    assert all(r[2] == 5 for r in res), res


def test_single_process_is_a_bypass():
    from vllm_ltr_b200 import step_broadcast as sb
    port = _free_port()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        d = {"a": torch.arange(3)}
        assert sb.broadcast_tensor_dict(d) is d
    finally:
        dist.destroy_process_group()
