#!/usr/bin/env python
"""Multi-GPU check + timing of the fused tensor-parallel exchange (run under torchrun on a GPU box):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_check.py

Parity: integer-valued inputs so the cross-rank sum is exact (like the reference's
tests/distributed/test_custom_all_reduce.py:30-51) against NCCL all_reduce + the oracle's
fused_add_rms_norm; then device time of fused vs NCCL all_reduce + our fused_add_rms_norm."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from oracle import kernels as K
    from vllm_ltr_b200 import _C
    from vllm_ltr_b200.parallel import FusedTPComm
    ok = True
    for dtype in (torch.bfloat16, torch.float16):
        for T, H in ((1, 4096), (7, 4096), (256, 4096), (256, 8192), (768, 8192), (300, 5120)):
            # both epoch modes: counted by the kernels (default, graph-safe) / passed by the host
            comm = FusedTPComm(768, H, dtype, dev, device_epoch=(T != 7))
            g = torch.Generator(device=dev).manual_seed(100 + rank)
            x = torch.randint(-4, 5, (T, H), generator=g, device=dev).to(dtype)
            g0 = torch.Generator(device=dev).manual_seed(7)
            resid = torch.randint(-4, 5, (T, H), generator=g0, device=dev).to(dtype)
            w = (torch.randn(H, generator=g0, device=dev) * 0.1 + 1).to(dtype)
            # reference path: NCCL sum, then the oracle's fused_add_rms_norm
            s = x.clone()
            dist.all_reduce(s)
            want_out, want_res = K.fused_add_rms_norm(s, resid, w, 1e-5)
            comm.x[0][:T].copy_(x)
            comm.resid[:T].copy_(resid)
            torch.cuda.synchronize()
            dist.barrier()
            out = comm.allreduce_add_rms_norm(0, w, 1e-5, T)
            torch.cuda.synchronize()
            e_res = torch.equal(comm.resid[:T], want_res)
            diff = (out.float() - want_out.float()).abs().max().item()
            e_out = diff <= 2 * float(want_out.float().abs().max()) * (2 ** -7 if dtype == torch.bfloat16 else 2 ** -10)
            comm.x[1][:T].copy_(x)
            torch.cuda.synchronize()
            dist.barrier()
            plain = comm.allreduce(1, T)
            torch.cuda.synchronize()
            e_plain = torch.equal(plain, s)
            if not (e_res and e_out and e_plain):
                ok = False
            if rank == 0:
                print(f"{dtype} T={T} H={H}: resid exact={e_res} out maxdiff={diff:.3g} ok={e_out} plain exact={e_plain}", flush=True)
            del comm
    # CUDA-graph replay of a chain of exchanges (device-side epochs): three replays must give
    # what three eager rounds give
    T, H, dtype = 256, 4096, torch.bfloat16
    comm = FusedTPComm(256, H, dtype, dev)
    g = torch.Generator(device=dev).manual_seed(200 + rank)
    xs = [torch.randint(-3, 4, (T, H), generator=g, device=dev).to(dtype) for _ in range(2)]
    w = torch.ones(H, device=dev, dtype=dtype)

    def chain():
        comm.allreduce_add_rms_norm(0, w, 1e-5, T)
        comm.allreduce_add_rms_norm(1, w, 1e-5, T)

    def reset():
        comm.x[0][:T].copy_(xs[0])
        comm.x[1][:T].copy_(xs[1])
        comm.resid[:T].zero_()
        torch.cuda.synchronize()
        dist.barrier()

    reset()
    for _ in range(3):
        chain()
    torch.cuda.synchronize()
    want_res, want_out = comm.resid[:T].clone(), comm.out[:T].clone()
    reset()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        chain()                                   # warm-up round (counts as round 1)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        chain()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    e_graph = torch.equal(comm.resid[:T], want_res) and torch.equal(comm.out[:T], want_out)
    ok = ok and e_graph
    if rank == 0:
        print(f"graph replay of 2 exchanges x 2 (after 1 eager round): exact={e_graph}", flush=True)
    del comm
    # timing at the decode shapes
    for T, H in ((256, 4096), (256, 8192), (768, 8192)):
        dtype = torch.bfloat16
        comm = FusedTPComm(768, H, dtype, dev)
        x = torch.randn(T, H, device=dev).to(dtype)
        resid = torch.randn(T, H, device=dev).to(dtype)
        w = torch.ones(H, device=dev, dtype=dtype)
        comm.x[0][:T].copy_(x)
        comm.resid[:T].copy_(resid)

        def fused():
            comm.allreduce_add_rms_norm(0, w, 1e-5, T)

        def nccl():
            dist.all_reduce(x)
            _C.ops.fused_add_rms_norm(x, resid, w, 1e-5)

        for name, fn in (("fused", fused), ("nccl+norm", nccl)):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 50 * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"T={T} H={H} world={world} {name}: {t.item():.1f} us per exchange", flush=True)
            x.copy_(torch.randn(T, H, device=dev).to(dtype))
    if rank == 0:
        print("TP CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
