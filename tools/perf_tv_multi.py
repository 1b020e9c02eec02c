#!/usr/bin/env python
"""configs[4] of BASELINE.json on N GPUs: TV denoising of one 8192 x 8192 fp32 image, row-sharded, consensus-form Douglas-Rachford
(K10).  Launch: torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/perf_tv_multi.py   (N = 1 works without torchrun).
Prints one JSON line on rank 0 -> paste into profiles/."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200.host import Context, DeviceExchangeComm  # noqa: E402


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
    side = 8192
    rows = side // world
    r0 = rank * rows
    gen = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randn(side, side, device="cuda", generator=gen)      # same image on every rank; each keeps its rows
    b = img[r0:r0 + rows].contiguous()
    del img
    f = pa.TVSplit(b, 0.3, comm=comm, row0=r0, Hglob=side) if world > 1 else pa.TVSplit(b, 0.3)
    it = pa.DouglasRachfordIteration(f.initial_point(), f=f, g=pa.IndConsensus(5), gamma=np.float32(1.0), comm=comm)
    st = it.step(None)
    for _ in range(20):
        st = it.step(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    K = 200
    t0 = time.perf_counter()
    for _ in range(K):
        st = it.step(st)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax)
    if rank == 0:
        nbytes = 11 * 4 * side * side
        print(json.dumps(dict(workload="TV denoising 8192x8192 fp32, consensus-form Douglas-Rachford, row-sharded", n_gpus=world, iterations=K,
                              ms_per_iteration=1e3 * dt / K, it_per_s=K / dt, aggregate_algorithmic_gbs=nbytes * K / dt / 1e9,
                              res_inf=float(st.res_norm_inf), halo="kernel reads one peer row over NVLink (cudaIpc)", exchange="device")), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
