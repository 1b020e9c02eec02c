#!/usr/bin/env python
"""BASELINE.json configs[2]-shaped dense least squares on N GPUs: A 10000 x 100000 fp32 (4 GB), COLUMN-sharded over the ranks
(SURVEY.md section 8e "Dense-A gradient under this partition"), IndBox / NormL1 + fixed-stepsize FastForwardBackward through the native
driver loop.  Per product the chunk partials of A x are all-gathered INSIDE the combine kernel over NVLink (csrc/lsq_kernels.cu:
k_gemv_n_combine_x) and folded in global chunk order, so f, every scalar and the iterates are bit-identical for every N.
Launch: torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/perf_dense_multi.py  (N = 1 without torchrun).
Strong scaling (the matrix is fixed); prints one JSON line on rank 0 -> gpurun_out/perf_dense_multi_n{N}.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, DeviceExchangeComm, dense_shard_bounds  # noqa: E402


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
    m, n = 10_000, 100_000
    lo, hi = dense_shard_bounds(np.float32, m, n, world)[rank]
    k = hi - lo
    # counter-based data (a function of the GLOBAL element index): every N works on the same matrix
    A_cm = torch.empty(k, m, device="cuda")               # (columns, rows) row-major == column-major shard
    L.check(ctx.lib.pb_fill_counter(ctx.h, L.PB_F32, k * m, lo * m, 11, 1.0 / np.sqrt(m), C.c_void_p(A_cm.data_ptr())))
    b = torch.empty(m, device="cuda")
    L.check(ctx.lib.pb_fill_counter(ctx.h, L.PB_F32, m, 0, 12, 1.0, C.c_void_p(b.data_ptr())))
    x = torch.empty(k, device="cuda")
    L.check(ctx.lib.pb_fill_counter(ctx.h, L.PB_F32, k, lo, 13, 1.0, C.c_void_p(x.data_ptr())))
    f = pa.LeastSquares(A_cm.t(), b, comm=comm, n_global=n, col_offset=lo) if world > 1 else pa.LeastSquares(A_cm.t(), b)
    del A_cm
    grad = torch.empty(k, device="cuda")
    # ---- the product pair alone: r = A x - b (+ in-kernel all-gather), grad = A_p' r
    for _ in range(3):
        val = f.value_and_gradient_into(ctx, x, grad)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        val = f.value_and_gradient_into(ctx, x, grad)
    e1.record()
    torch.cuda.synchronize()
    pair_ms = e0.elapsed_time(e1) / reps
    row = ctx.read_scalars()
    fval = float(val.resolve(row, None))
    gsum = float(grad.double().abs().sum())
    # ---- fixed-stepsize FISTA through the native driver loop
    K = 100
    solver = pa.FastForwardBackward(maxit=K, tol=-1.0)
    kw = dict(x0=torch.zeros(k, device="cuda"), f=f, g=pa.NormL1(0.05), Lf=20.0, comm=comm, n_global=n)
    solver(**kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    z, it = solver(**kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt, pair_ms], device="cuda", dtype=torch.float64)
    gs = torch.tensor([gsum], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(gs)
    dt, pair_ms = float(tt[0]), float(tt[1])
    st = solver.last_state
    if rank == 0:
        bytes_pair = 2.0 * m * n * 4 / world
        out = {"workload": "dense 10000 x 100000 fp32 least squares, column-sharded", "n_gpus": world, "columns_per_gpu": k,
               "pair_ms": pair_ms, "pair_gbs_per_gpu": bytes_pair / (pair_ms * 1e-3) / 1e9,
               "fista_iterations": int(it), "fista_ms_per_iteration": 1e3 * dt / it, "fista_it_per_s": it / dt, "driver": solver.last_driver,
               "parity": {"f_at_x": fval, "sum_abs_grad": float(gs[0]), "f_x_end": float(st.f_x), "g_z_end": float(st.g_z),
                          "res_inf_end": float(st.res_norm_inf)}}
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"perf_dense_multi_n{world}.json"), "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
