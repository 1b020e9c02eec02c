#!/usr/bin/env python
"""Timing of the "next" rows on one B200: the K7 / K8 kernels alone and the PANOC / Douglas-Rachford
iterations built from them, at the BASELINE.json configs[3] / configs[4] sizes.  Parity-test cases, not bench lines; the
numbers go to profiles/.  -> gpurun_out/perf_next.json

Algorithmic bytes (s = element size): lincomb2 3s, L-BFGS update 6s, one two-loop link 4s (read d, u, w; write d; the last
link +2s for x, x_d), the whole apply with memory m: 2s + 4s*2m + 2s, fused DR pass 2s (+1s with a data vector)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, ptr  # noqa: E402
from tune_step import timeit  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6451.8


def main():
    PEAK = peak()
    ctx = Context.get()
    out = {"peak_gbs": PEAK}
    quick = "--quick" in sys.argv
    ngroups, gsz = (100_000, 128) if quick else (1_000_000, 128)
    n = ngroups * gsz                                      # configs[3]: 1e6 groups x 128 fp32
    es = 4

    def rec(name, ms, nbytes, **kw):
        out[name] = dict(ms=ms, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK, algorithmic_bytes=nbytes, **kw)
        print(name, out[name], flush=True)

    x, y = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    o1, o2 = torch.empty_like(x), torch.empty_like(x)
    rec("k7_lincomb2", timeit(lambda: L.check(ctx.lib.pb_lincomb2(ctx.h, L.PB_F32, n, 0.25, ptr(x), 0.75, ptr(y), ptr(o1))), reps=20), 3 * es * n)
    H = pa.LBFGS(5).initialize(x)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(6):                                     # fill the ring with positive-curvature pairs
        s = torch.randn(n, device="cuda", generator=gen)
        yv = 0.5 * s + 0.1 * torch.randn(n, device="cuda", generator=gen)
        assert H.update(s, yv)
    del s, yv
    rec("k7_lbfgs_update", timeit(lambda: H.enqueue_update(x, y, o1, o2), reps=20), 6 * es * n)
    H.commit(pa.Scalars(ctx.read_scalars()[None, :]))
    m = H.currmem
    d, xd = torch.empty_like(x), torch.empty_like(x)
    l0 = ctx.launches()
    ms = timeit(lambda: H.mul_into(d, y, scale=-1.0, x=x, x_d=xd), reps=10)
    nb = (2 + 4 * 2 * m + 2) * es * n
    rec("k7_lbfgs_apply_m5", ms, nb, launches_per_apply=(ctx.launches() - l0) // 13, currmem=m)
    # the same fused step PANOC uses: K1 with res materialised, NormL21
    desc = L.pb_prox(L.PB_PROX_L21, gsz, 0.5, 0.0, None, None)
    rec("k1_fb_step_l21_with_res", timeit(lambda: L.check(ctx.lib.pb_fb_step(ctx.h, L.PB_F32, n, ptr(x), ptr(y), 0.1, C.byref(desc), None, ptr(o1), ptr(o2))), reps=20), 4 * es * n)
    del H, d, xd, o1, o2, x, y
    torch.cuda.empty_cache()

    # ---- configs[3]: Group-Lasso, PANOC + LBFGS(5) + NormL21, f = 0.5||Ax - b||^2 with A = blockdiag (nblk blocks of mb x nb) ----
    nblk, mb = (100, 4) if quick else (1000, 4)
    nb_ = n // nblk
    A = torch.randn(nblk, nb_, mb, device="cuda", generator=gen) / np.sqrt(nb_)
    xt = torch.zeros(ngroups, gsz, device="cuda")
    act = torch.randperm(ngroups, device="cuda", generator=gen)[: ngroups // 100]
    xt[act] = torch.randn(act.numel(), gsz, device="cuda", generator=gen)
    xt = xt.view(-1)
    f0 = pa.BlockDiagLeastSquares(A, torch.zeros(nblk * mb, device="cuda"))
    tmp = torch.empty_like(xt)
    f0.value_and_gradient_into(ctx, xt, tmp)
    b = f0.r.clone() + 0.01 * torch.randn(nblk * mb, device="cuda", generator=gen)
    f = pa.BlockDiagLeastSquares(A, b)
    f.value_and_gradient_into(ctx, torch.zeros_like(xt), tmp)
    lam = 0.1 * float(tmp.view(ngroups, gsz).norm(dim=1).max())
    del f0, tmp, xt
    torch.cuda.empty_cache()
    xg, gg = torch.randn(n, device="cuda", generator=gen), torch.empty(n, device="cuda")
    rec("k4_blockdiag_residual_4x128000", timeit(lambda: L.check(ctx.lib.pb_lsq_blockdiag_residual(ctx.h, L.PB_F32, nblk, mb, nb_, ptr(A), ptr(xg), ptr(b), ptr(f.r))), reps=20), A.numel() * es + n * es)
    rec("k4_blockdiag_gradient_4x128000", timeit(lambda: L.check(ctx.lib.pb_lsq_blockdiag_gradient(ctx.h, L.PB_F32, nblk, mb, nb_, ptr(A), ptr(f.r), ptr(gg))), reps=20), A.numel() * es + n * es)
    del xg, gg
    for K, label in ((30, "config3_panoc_lbfgs5_l21"),):
        it = pa.PANOCIteration(torch.zeros(n, device="cuda"), f=f, g=pa.NormL21(lam, gsz))
        st = it.init()
        for _ in range(8):                                 # warm-up: fills the L-BFGS memory
            st = it.step(st)
        torch.cuda.synchronize()
        l0, t0 = ctx.launches(), time.perf_counter()
        for _ in range(K):
            st = it.step(st)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = (ctx.launches() - l0) / K
        mem = st.H.currmem
        # traffic of one accepted iteration (fixed gamma would skip the 2 A-passes of the stepsize test): apply chain, f at x_d (2 passes
        # over A), fused step with res, update; adaptive: + f at z (2 passes over A)
        vec = (2 + 8 * mem + 2) + 4 + 6
        nbytes = vec * es * n + 4 * A.numel() * es
        out[label] = dict(iterations=K, ms_per_iteration=1e3 * dt / K, it_per_s=K / dt, launches_per_iteration=launches, currmem=mem,
                          tau_backtracks=it.tau_backtracks, gamma_backtracks=it.backtracks, algorithmic_bytes_per_iteration=nbytes,
                          gbs=nbytes * K / dt / 1e9, frac=nbytes * K / dt / 1e9 / PEAK, res_inf_over_gamma=float(st.res_norm_inf / st.gamma),
                          n=n, groups=ngroups, A_bytes=A.numel() * es, lam=lam)
        print(label, out[label], flush=True)
    del it, st, f, A
    torch.cuda.empty_cache()

    # ---- configs[4] shape: 8192 x 8192 fp32 image, fused Douglas-Rachford pass (f = 0.5||x - b||^2, g = NormL1: element-wise stand-in) ----
    side = 2048 if quick else 8192
    npx = side * side
    bimg, x0 = torch.randn(npx, device="cuda"), torch.randn(npx, device="cuda")
    fd = pa.SqrNormL2(1.0, bimg).descriptor(np.float32)
    gd = pa.NormL1(0.3).descriptor(np.float32)
    x1 = torch.empty_like(x0)
    rec("k8_dr_step_8192sq", timeit(lambda: L.check(ctx.lib.pb_dr_step(ctx.h, L.PB_F32, npx, ptr(x0), 0.7, C.byref(fd), C.byref(gd), ptr(x1), None, None, None, None)), reps=50), 3 * es * npx)
    sweep = []
    for unroll in (1, 2, 4):
        for ctas in (1, 2, 3, 4, 6, 8):
            ctx.set_launch(ctas_per_sm=ctas, unroll=unroll)
            ms = timeit(lambda: L.check(ctx.lib.pb_dr_step(ctx.h, L.PB_F32, npx, ptr(x0), 0.7, C.byref(fd), C.byref(gd), ptr(x1), None, None, None, None)), reps=30)
            sweep.append(dict(unroll=unroll, ctas_per_sm=ctas, ms=ms, frac=3 * es * npx / ms / 1e6 / PEAK))
    ctx.set_launch()
    out["k8_dr_step_sweep"] = sweep
    print("k8 sweep", sorted(sweep, key=lambda r: r["ms"])[:6], flush=True)
    gd2 = pa.IndBox(-0.5, 0.5).descriptor(np.float32)
    fd2 = pa.NormL1(0.3).descriptor(np.float32)
    rec("k8_dr_step_8192sq_l1_box_inplace", timeit(lambda: L.check(ctx.lib.pb_dr_step(ctx.h, L.PB_F32, npx, ptr(x0), 0.7, C.byref(fd2), C.byref(gd2), ptr(x0), None, None, None, None)), reps=50), 2 * es * npx)
    K = 200
    alg = pa.DouglasRachford(maxit=K, tol=-1.0)
    alg(x0=x0, f=pa.SqrNormL2(1.0, bimg), g=pa.NormL1(0.3), gamma=0.7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    yv, k = alg(x0=x0, f=pa.SqrNormL2(1.0, bimg), g=pa.NormL1(0.3), gamma=0.7)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["config4shape_dr_solver_elementwise"] = dict(iterations=k, ms_per_iteration=1e3 * dt / k, it_per_s=k / dt, gbs=3 * es * npx * k / dt / 1e9,
                                                     frac=3 * es * npx * k / dt / 1e9 / PEAK, note="includes x0 copy, per-iteration scalar read-back, final y materialisation")
    print(out["config4shape_dr_solver_elementwise"], flush=True)
    del x0, x1, bimg, yv
    torch.cuda.empty_cache()
    # ---- configs[4]: TV denoising of the 8192 x 8192 image, consensus-form Douglas-Rachford, one fused pass per iteration (K10) ----
    bt = torch.randn(side, side, device="cuda")
    ftv = pa.TVSplit(bt, 0.3)
    X0 = ftv.initial_point()
    X1 = torch.empty_like(X0)
    rec("k10_dr_tv_step_8192sq", timeit(lambda: L.check(ctx.lib.pb_dr_tv_step(ctx.h, L.PB_F32, side, side, ptr(X0), ptr(ftv.b), 1.0, 0.3, ptr(X1), None, None, 0, side, None, None)), reps=20), 11 * es * npx)
    for ctas in (2, 3, 4):
        ctx.set_launch(ctas_per_sm=ctas)
        ms = timeit(lambda: L.check(ctx.lib.pb_dr_tv_step(ctx.h, L.PB_F32, side, side, ptr(X0), ptr(ftv.b), 1.0, 0.3, ptr(X1), None, None, 0, side, None, None)), reps=10)
        print("k10 ctas_per_sm", ctas, ms, 11 * es * npx / ms / 1e6 / PEAK, flush=True)
        out.setdefault("k10_sweep", []).append(dict(ctas_per_sm=ctas, ms=ms, frac=11 * es * npx / ms / 1e6 / PEAK))
    ctx.set_launch()
    del X1
    K = 100
    alg = pa.DouglasRachford(maxit=K, tol=-1.0)
    alg(x0=X0, f=ftv, g=pa.IndConsensus(5), gamma=1.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    yv, k = alg(x0=X0, f=ftv, g=pa.IndConsensus(5), gamma=1.0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["config4_tv_dr_solver"] = dict(iterations=k, ms_per_iteration=1e3 * dt / k, it_per_s=k / dt, gbs=11 * es * npx * k / dt / 1e9,
                                       frac=11 * es * npx * k / dt / 1e9 / PEAK, note="includes x0 copy, per-iteration scalar read-back, final y/z materialisation")
    print(out["config4_tv_dr_solver"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "perf_next.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
