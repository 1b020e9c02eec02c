#!/usr/bin/env python
"""The reference's benchmark suite (benchmark/benchmarks.jl:47-93: ForwardBackward, FastForwardBackward, PANOC, DouglasRachford on
the three Lasso fixtures, Float64, the suite's own settings) timed on the B200 path and on the numpy oracle beside it.
    python tools/perf_fixtures.py            # GPU + oracle  -> gpurun_out/perf_fixtures.json
    python tools/perf_fixtures.py --cpu-only # oracle only (no GPU needed)
These problems are tiny (5x10 .. 500x1000): they measure launch / synchronisation latency, not bandwidth; the numbers are reported so
that nobody has to guess who wins where."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fb_oracle as o  # noqa: E402
from oracle import panoc_oracle as po  # noqa: E402


def best_of(fn, reps=3):
    out, best = None, float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return out, best


def main():
    cpu_only = "--cpu-only" in sys.argv
    pa = None
    if not cpu_only:
        import proxb200 as pa_  # noqa: E402

        pa = pa_
    res = {}

    def config0():
        # BASELINE.json configs[0] as written: 200 x 500 dense fp64 Lasso (seed 1, 25 non-zeros, lambda = 0.1 ||A'b||_inf)
        rng = np.random.default_rng(1)
        A_ = rng.standard_normal((200, 500))
        xt = np.zeros(500)
        xt[rng.choice(500, 25, replace=False)] = rng.standard_normal(25)
        b_ = A_ @ xt + 0.01 * rng.standard_normal(200)
        return dict(A=A_, b=b_, lam=0.1 * np.max(np.abs(A_.T @ b_)))

    for name in ("tiny", "small", "medium", "config0_200x500"):
        d = config0() if name.startswith("config0") else np.load(os.path.join(ROOT, "tests", "golden", f"lasso_{name}.npz"))
        A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
        n = A.shape[1]
        x0 = np.zeros(n)
        # the terms are built once, outside the timed region -- as the reference's suite does (`setup = (f = LeastSquares($A, $b) ...)`,
        # benchmark/benchmarks.jl:49-53): for the GPU arm this keeps the one-off upload of A out of the solve time
        fo, go, fo_sq, fo_px = o.LeastSquares(A, b), o.NormL1(lam), o.SquaredDistance(b), po.LeastSquaresProx(A, b)
        if pa is not None:
            fg, gg, fg_sq = pa.LeastSquares(A, b), pa.NormL1(lam), pa.SquaredDistance(b)
        cases = {
            "ForwardBackward": (lambda: o.forward_backward(x0, fo, go, tol=1e-6), lambda s: s(x0=x0, f=fg, g=gg), lambda: pa.ForwardBackward(tol=1e-6)),
            "FastForwardBackward": (lambda: o.fast_forward_backward(x0, fo, go, tol=1e-6), lambda s: s(x0=x0, f=fg, g=gg), lambda: pa.FastForwardBackward(tol=1e-6)),
            "PANOC": (lambda: po.panoc(x0, f=fo_sq, A=A, g=go, tol=1e-6), lambda s: s(x0=x0, f=fg_sq, A=A, g=gg), lambda: pa.PANOC(tol=1e-6)),
            "DouglasRachford": (lambda: po.douglas_rachford(x0, f=fo_px, g=go, gamma=1.0, tol=1e-6), lambda s: s(x0=x0, f=fg, g=gg, gamma=1.0),
                                lambda: pa.DouglasRachford(tol=1e-6)),
        }
        if "--fb-only" in sys.argv:
            cases = {k: v for k, v in cases.items() if k.endswith("ForwardBackward")}
        for alg, (f_cpu, f_gpu, mk) in cases.items():
            (_, it_c), t_c = best_of(f_cpu)
            row = {"oracle_iterations": int(it_c), "oracle_seconds": t_c, "oracle_it_per_s": it_c / t_c}
            if pa is not None:
                solver = mk()
                f_gpu(solver)                            # warm-up (allocations, module load)
                (_, it_g), t_g = best_of(lambda: f_gpu(solver), reps=5)
                row.update(gpu_iterations=int(it_g), gpu_seconds=t_g, gpu_it_per_s=it_g / t_g)
                if alg in ("ForwardBackward", "FastForwardBackward"):
                    # device time of the solve alone (CUDA events around the persistent kernel) and where CTA 0 spent its cycles
                    import ctypes as C

                    from proxb200 import _lib as L
                    from proxb200.host import Context

                    solver.profile = True
                    f_gpu(solver)
                    tm = getattr(solver, "last_timing", {})
                    row.update(persistent_ctas=int(getattr(solver, "last_persistent_ctas", 0)), gpu_kernel_ms=tm.get("loop_ms"),
                               gpu_kernel_it_per_s=(it_g / (tm["loop_ms"] * 1e-3)) if tm.get("loop_ms") else None)
                    cyc = (C.c_int64 * 8)()
                    ctx = Context.get()
                    L.check(ctx.lib.pb_persist_phase_cycles(ctx.h, cyc))
                    names = ["gemv_n", "barrier_fold", "combine", "gemv_t", "step", "other"]
                    row["cycles_per_iteration"] = {nm: round(cyc[i] / max(1, it_g), 1) for i, nm in enumerate(names)}
                    solver.profile = False
            res[f"{name}/{alg}"] = row
            print(f"{name}/{alg}", row, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "perf_fixtures_cpu.json" if cpu_only else "perf_fixtures.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
