#!/usr/bin/env python
"""Time the least-squares kernels (dense + block-diagonal) and whole solves on the fixtures.  -> gpurun_out/tune_lsq.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200.host import Context  # noqa: E402
from tune_step import timeit  # noqa: E402


def main():
    ctx = Context.get()
    res = []
    # block-diagonal, cfg2 shape: 100 blocks of 100 x 1e5 fp32 (4 GB)
    for nblk, mb, nb in ((100, 100, 100_000), (1000, 100, 10_000), (100, 128, 100_000), (10000, 16, 1024)):
        A = torch.randn(nblk, nb, mb, device="cuda") * 0.1
        b = torch.randn(nblk * mb, device="cuda")
        x = torch.randn(nblk * nb, device="cuda")
        f = pa.BlockDiagLeastSquares(A, b)
        grad = torch.empty_like(x)
        ms_r = timeit(lambda: f.value_into(ctx, x), reps=5)
        ms_f = timeit(lambda: f.value_and_gradient_into(ctx, x, grad), reps=5)
        byt = A.numel() * 4
        res.append(dict(kind="blockdiag", nblk=nblk, mb=mb, nb=nb, ms_residual=ms_r, ms_value_and_gradient=ms_f,
                        gbs_residual=byt / ms_r / 1e6, gbs_both=2 * byt / ms_f / 1e6))
        print(res[-1], flush=True)
        del A, b, x, f, grad
        torch.cuda.empty_cache()
    # dense
    for m, n, dt in ((500, 1000, torch.float64), (200, 500, torch.float64), (10_000, 100_000, torch.float32), (50, 100, torch.float64)):
        A = torch.randn(m, n, device="cuda", dtype=dt)
        b = torch.randn(m, device="cuda", dtype=dt)
        x = torch.randn(n, device="cuda", dtype=dt)
        f = pa.LeastSquares(A, b)
        grad = torch.empty_like(x)
        ms_f = timeit(lambda: f.value_and_gradient_into(ctx, x, grad), reps=20)
        byt = A.numel() * A.element_size()
        res.append(dict(kind="dense", m=m, n=n, dtype=str(dt), ms_value_and_gradient=ms_f, gbs_both=2 * byt / ms_f / 1e6))
        print(res[-1], flush=True)
        del A, f
    # The tensor-core question: would the 2-right-hand-side form A*[z x] (SURVEY.md section 7.8 ii) on the tensor pipe beat the
    # streaming GEMV?  Library GEMM (cuBLAS through torch, TF32 tensor cores allowed) on the same 4 GB matrix with 1, 2 and 8 right-hand sides,
    # against our single-RHS product: all of them are bound by reading A once from HBM (off the parity path: different arithmetic).
    m, n = 10_000, 100_000
    A = torch.randn(m, n, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    for k in (1, 2, 8):
        X = torch.randn(n, k, device="cuda")
        ms = timeit(lambda: torch.mm(A, X), reps=20)
        res.append(dict(kind="cublas_tf32_gemm_A_times_k_rhs", m=m, n=n, k=k, ms=ms, gbs_of_A=A.numel() * 4 / ms / 1e6))
        print(res[-1], flush=True)
    b = torch.randn(m, device="cuda")
    x = torch.randn(n, device="cuda")
    f = pa.LeastSquares(A, b)
    ms = timeit(lambda: f.value_into(ctx, x), reps=20)
    res.append(dict(kind="ours_residual_single_rhs", m=m, n=n, ms=ms, gbs_of_A=A.numel() * 4 / ms / 1e6))
    print(res[-1], flush=True)
    del A, f
    torch.cuda.empty_cache()
    # whole solves on the reference fixtures (adaptive, tol 1e-6) -> iterations/s
    for name in ("tiny", "small", "medium"):
        d = np.load(os.path.join(ROOT, "tests", "golden", f"lasso_{name}.npz"))
        for alg in ("ffb", "fb"):
            solver = (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=1e-6)
            f = pa.LeastSquares(d["A"], d["b"])
            solver(x0=np.zeros(d["A"].shape[1]), f=f, g=pa.NormL1(float(d["lam"])))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            z, it = solver(x0=np.zeros(d["A"].shape[1]), f=f, g=pa.NormL1(float(d["lam"])))
            dt_ = time.perf_counter() - t0
            res.append(dict(kind="solve", fixture=name, alg=alg, iterations=it, seconds=dt_, it_per_s=it / dt_))
            print(res[-1], flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_lsq.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
