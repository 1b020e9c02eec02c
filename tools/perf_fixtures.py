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
    for name in ("tiny", "small", "medium"):
        d = np.load(os.path.join(ROOT, "tests", "golden", f"lasso_{name}.npz"))
        A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
        n = A.shape[1]
        x0 = np.zeros(n)
        cases = {
            "ForwardBackward": (lambda: o.forward_backward(x0, o.LeastSquares(A, b), o.NormL1(lam), tol=1e-6),
                                lambda: pa.ForwardBackward(tol=1e-6)(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam))),
            "FastForwardBackward": (lambda: o.fast_forward_backward(x0, o.LeastSquares(A, b), o.NormL1(lam), tol=1e-6),
                                    lambda: pa.FastForwardBackward(tol=1e-6)(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam))),
            "PANOC": (lambda: po.panoc(x0, f=o.SquaredDistance(b), A=A, g=o.NormL1(lam), tol=1e-6),
                      lambda: pa.PANOC(tol=1e-6)(x0=x0, f=pa.SquaredDistance(b), A=A, g=pa.NormL1(lam))),
            "DouglasRachford": (lambda: po.douglas_rachford(x0, f=po.LeastSquaresProx(A, b), g=o.NormL1(lam), gamma=1.0, tol=1e-6),
                                lambda: pa.DouglasRachford(tol=1e-6)(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), gamma=1.0)),
        }
        for alg, (f_cpu, f_gpu) in cases.items():
            (_, it_c), t_c = best_of(f_cpu)
            row = {"oracle_iterations": int(it_c), "oracle_seconds": t_c, "oracle_it_per_s": it_c / t_c}
            if pa is not None:
                f_gpu()                                  # warm-up (allocations, module load)
                (_, it_g), t_g = best_of(f_gpu)
                row.update(gpu_iterations=int(it_g), gpu_seconds=t_g, gpu_it_per_s=it_g / t_g)
            res[f"{name}/{alg}"] = row
            print(f"{name}/{alg}", row, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "perf_fixtures_cpu.json" if cpu_only else "perf_fixtures.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
