#!/usr/bin/env python
"""Timing of the other BASELINE.json configs on one B200 (they are parity-test cases, not bench lines; numbers go to
profiles/).  -> gpurun_out/perf_configs.json"""
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
from proxb200.host import Context, DeviceExchangeComm, ptr  # noqa: E402
from tune_step import timeit  # noqa: E402

PEAK = 6451.8   # MEASURED_PEAKS.json hbm_gbs of this pool


def main():
    ctx = Context.get()
    out = {}
    # configs[1]: Lasso, A block-diagonal 100 x (100 x 1e5) fp32 (n = 1e7, m = 1e4), FISTA + NormL1, fixed gamma
    nblk, mb, nb = 100, 100, 100_000
    gen = torch.Generator(device="cuda").manual_seed(2)
    A = torch.randn(nblk, nb, mb, device="cuda", generator=gen) / 10.0
    xt = torch.zeros(nblk * nb, device="cuda")
    idx = torch.randint(0, nblk * nb, (10_000,), device="cuda", generator=gen)
    xt[idx] = torch.randn(10_000, device="cuda", generator=gen)
    f0 = pa.BlockDiagLeastSquares(A, torch.zeros(nblk * mb, device="cuda"))
    g0 = torch.empty_like(xt)
    f0.value_and_gradient_into(ctx, xt, g0)
    b = f0.r.clone() + 0.01 * torch.randn(nblk * mb, device="cuda", generator=gen)
    f = pa.BlockDiagLeastSquares(A, b)
    gb = torch.empty_like(xt)
    f.value_and_gradient_into(ctx, torch.zeros_like(xt), gb)
    lam = float(0.1 * gb.abs().max())
    v = torch.randn_like(xt)
    for _ in range(20):      # power iterations for L = ||A||^2
        fz = pa.BlockDiagLeastSquares(A, torch.zeros_like(b))
        w = torch.empty_like(v)
        fz.value_and_gradient_into(ctx, v, w)
        Lhat = float(w.norm() / v.norm())
        v = w / w.norm()
    comm = DeviceExchangeComm(ctx)
    try:
        K = 200
        x0 = torch.zeros_like(xt)
        for mode, key in ((-1, "config1_blockdiag_fista_n1e7_two_sweeps"), (0, "config1_blockdiag_fista_n1e7")):
            # mode -1: residual + gradient + fused step kernels (A swept twice per iteration); 0 = auto: csrc/lsq_fista.cu (A swept once)
            L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_LSQ_FISTA, mode))
            solver = pa.FastForwardBackward(maxit=K, tol=-1.0)
            solver(x0=x0, f=f, g=pa.NormL1(lam), Lf=1.05 * Lhat, comm=comm)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            z, it = solver(x0=x0, f=f, g=pa.NormL1(lam), Lf=1.05 * Lhat, comm=comm)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st = solver.last_state
            sweeps = 2 if mode < 0 else 1
            bytes_iter = sweeps * A.numel() * 4 + 5 * 4 * xt.numel() + 4 * 4 * b.numel()
            out[key] = dict(iterations=it, seconds=dt, it_per_s=it / dt, ms_per_iteration=1e3 * dt / it, sweeps_of_A_per_iteration=sweeps,
                            algorithmic_bytes_per_iteration=bytes_iter, gbs=bytes_iter * it / dt / 1e9,
                            frac_of_measured_peak=bytes_iter * it / dt / 1e9 / PEAK, Lhat=Lhat, lam=lam,
                            final_res_inf_over_gamma=float(st.res_norm_inf / st.gamma), nnz=int((z != 0).sum()), parity=dict(solver.last_parity))
            print(key, out[key], flush=True)
        L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_LSQ_FISTA, 0))
        # same problem, adaptive (backtracking) FISTA to tol 1e-4: a real solve
        t0 = time.perf_counter()
        z2, it2 = pa.FastForwardBackward(maxit=3000, tol=1e-4)(x0=x0, f=f, g=pa.NormL1(lam), comm=comm)
        torch.cuda.synchronize()
        dt2 = time.perf_counter() - t0
        out["config1_blockdiag_fista_adaptive_solve"] = dict(iterations=it2, seconds=dt2, it_per_s=it2 / dt2)
        print(out["config1_blockdiag_fista_adaptive_solve"], flush=True)
    finally:
        comm.close()
    del A, f, f0, fz
    torch.cuda.empty_cache()
    # configs[2]: box-constrained, n = 1e8 fp32, ForwardBackward step K1 ("fused step only")
    n = 100_000_000
    x, g = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    z = torch.empty_like(x)
    desc = L.pb_prox(L.PB_PROX_BOX, 0, -1.0, 1.0, None, None)
    ms = timeit(lambda: L.check(ctx.lib.pb_fb_step(ctx.h, L.PB_F32, n, ptr(x), ptr(g), 0.1, C.byref(desc), None, ptr(z), None)), reps=50, warm=5)
    out["config2_box_fb_step_n1e8"] = dict(ms=ms, gbs=12 * n / ms / 1e6, frac=12 * n / ms / 1e6 / PEAK, steps_per_s=1e3 / ms)
    print(out["config2_box_fb_step_n1e8"], flush=True)
    # configs[3] shape: NormL21, 781250 groups x 128 fp32 (n = 1e8): fused step K2
    zp, xn = torch.randn(n, device="cuda"), torch.empty_like(x)
    desc = L.pb_prox(L.PB_PROX_L21, 128, 5.0, 0.0, None, None)
    ms = timeit(lambda: L.check(ctx.lib.pb_ffb_step(ctx.h, L.PB_F32, n, ptr(x), ptr(g), ptr(zp), 0.1, 0.3, C.byref(desc), None, ptr(z), None, ptr(xn))), reps=30, warm=5)
    out["config3_l21_ffb_step_n1e8"] = dict(ms=ms, gbs=20 * n / ms / 1e6, frac=20 * n / ms / 1e6 / PEAK)
    print(out["config3_l21_ffb_step_n1e8"], flush=True)
    ms = timeit(lambda: L.check(ctx.lib.pb_fb_step(ctx.h, L.PB_F32, n, ptr(x), ptr(g), 0.1, C.byref(desc), None, ptr(z), None)), reps=30, warm=5)
    out["config3_l21_fb_step_n1e8"] = dict(ms=ms, gbs=12 * n / ms / 1e6, frac=12 * n / ms / 1e6 / PEAK)
    print(out["config3_l21_fb_step_n1e8"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "perf_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
