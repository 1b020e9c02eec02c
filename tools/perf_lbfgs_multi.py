#!/usr/bin/env python
"""Micro-benchmark of the sharded L-BFGS apply (two-loop chain with in-kernel rank-combined dots) and of the fused step + exchange on N GPUs.
torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/perf_lbfgs_multi.py"""
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, DeviceExchangeComm, ptr  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context.get()
    comm = DeviceExchangeComm(ctx)
    n = 128_000_000 // world
    gen = torch.Generator(device="cuda").manual_seed(1 + rank)
    x, y = torch.randn(n, device="cuda", generator=gen), torch.randn(n, device="cuda", generator=gen)
    H = pa.LBFGS(5).initialize(x, comm=comm)
    for _ in range(6):
        s = torch.randn(n, device="cuda", generator=gen)
        yv = 0.5 * s + 0.1 * torch.randn(n, device="cuda", generator=gen)
        assert H.update(s, yv)
    del s, yv
    d, xd = torch.empty_like(x), torch.empty_like(x)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    out = {"n_gpus": world, "n_per_gpu": n}
    out["lbfgs_apply_ms"] = timed(lambda: H.mul_into(d, y, scale=-1.0, x=x, x_d=xd))
    desc = L.pb_prox(L.PB_PROX_L21, 128, 0.05, 0.0, None, None)
    z = torch.empty_like(x)

    def step():
        L.check(ctx.lib.pb_fb_step(ctx.h, L.PB_F32, n, ptr(x), ptr(y), 0.1, C.byref(desc), None, ptr(z), None))
        comm.exchange(ctx)

    out["l21_step_plus_exchange_ms"] = timed(step)
    out["exchange_only_ms"] = timed(lambda: (L.check(ctx.lib.pb_exchange(ctx.h)), comm.exchange(ctx)), reps=50)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
