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
    kw = dict(f=f, g=pa.NormL21(0.05, gsz), comm=comm, n_global=n, Lf=float(4.0))
    # time the ITERATIONS (panoc.jl:138-255), not the set-up: init allocates ~30 n-vectors and the 6 GB L-BFGS ring (cudaMalloc / cudaFree of
    # that size cost 10-700 ms and vary from call to call)
    itr = pa.PANOCIteration(x0, **kw)
    st = itr.init()
    for _ in range(8):
        st = itr.step(st)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        st = itr.step(st)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it = K
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])

    class _Alg:
        last_state = st
        last_iteration = itr

    alg = _Alg

    def run_solver():
        a_ = pa.PANOC(tol=-1.0, maxit=K)
        return a_(x0=x0, **kw)

    if "--trace" in sys.argv:
        # where does an iteration go?  every library call followed by a stream synchronisation, wall clock per entry point
        import collections

        from proxb200 import _lib as L_

        acc, cnt = collections.Counter(), collections.Counter()
        lib = ctx.lib

        class Traced:
            def __getattr__(self, name):
                fn = getattr(lib, name)
                if not name.startswith("pb_") or name in ("pb_last_error",):
                    return fn

                def wrapped(*a, **k):
                    t1 = time.perf_counter()
                    rc = fn(*a, **k)
                    torch.cuda.synchronize()
                    acc[name] += time.perf_counter() - t1
                    cnt[name] += 1
                    return rc

                return wrapped

        ctx.lib = Traced()
        t1 = time.perf_counter()
        run_solver()
        tot = time.perf_counter() - t1
        ctx.lib = lib
        if rank == 0:
            print("traced total ms/iteration", 1e3 * tot / K)
            for name, v in acc.most_common(14):
                print(f"  {name:40s} {1e3 * v / K:8.3f} ms/it  {cnt[name] / K:6.1f} calls/it")
        _ = L_
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
