#!/usr/bin/env python
"""configs[3] of BASELINE.json on N GPUs: Group-Lasso (NormL21, 1e6 groups x 128 fp32, n = 1.28e8), PANOC + LBFGS(5), f = 0.5||Ax - b||^2 with a
block-diagonal A (1000 blocks of 4 x 128000), row-sharded by whole blocks; every dot product of the L-BFGS recursion is summed over the
ranks inside its kernel.  Launch: torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/perf_panoc_multi.py  (N = 1 without torchrun).
Strong scaling (the problem is fixed); prints one JSON line on rank 0."""
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
    ngroups, gsz, nblk, mb = 1_000_000, 128, 1000, 4
    n = ngroups * gsz
    nb = n // nblk
    per = nblk // world
    k0 = rank * per
    # every block is generated from its own seed, so each rank builds exactly its blocks of the same global problem
    A = torch.empty(per, nb, mb, device="cuda")
    b = torch.empty(per * mb, device="cuda")
    for k in range(per):
        g = torch.Generator(device="cuda").manual_seed(1000 + k0 + k)
        A[k] = torch.randn(nb, mb, device="cuda", generator=g) / np.sqrt(nb)
        b[k * mb:(k + 1) * mb] = torch.randn(mb, device="cuda", generator=g)
    f = pa.BlockDiagLeastSquares(A, b, comm=comm)
    x0 = torch.zeros(per * nb, device="cuda")
    K = 30
    alg = pa.PANOC(tol=-1.0, maxit=K)
    kw = dict(x0=x0, f=f, g=pa.NormL21(0.05, gsz), comm=comm, n_global=n, Lf=float(4.0))
    alg(**kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    z, it = alg(**kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])
    if rank == 0:
        st = alg.last_state
        print(json.dumps(dict(workload="configs[3]: group lasso 1e6 x 128 fp32, PANOC + LBFGS(5) + NormL21, block-diagonal A, row shards", n_gpus=world,
                              iterations=it, ms_per_iteration=1e3 * dt / it, it_per_s=it / dt, res_inf=float(st.res_norm_inf), gamma=float(st.gamma),
                              tau_backtracks=alg.last_iteration.tau_backtracks)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
