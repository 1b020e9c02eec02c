#!/usr/bin/env python
"""PANOC on the reference's benchmark fixtures (benchmark/benchmarks.jl:71-77: f = SquaredDistance(b), A = A, g = NormL1, Float64), GPU
host vs numpy oracle STATE BY STATE for the whole solve: where do the two runs first take a different decision (stepsize gamma or
line-search tau), and how far apart were the scalars that decided it?  -> gpurun_out/panoc_divergence.json, profiles/r02_panoc_divergence.md
Background: PANOC's accept test (panoc.jl:196-205) compares FBE_new with FBE_x - sigma*||res||^2 + 10 eps (1 + |FBE_x|); near convergence
both sides agree to ~1e-15 relative, so last-ulp differences in the dot products of the L-BFGS recursion (BLAS order in the oracle and
in the Julia reference, exactly rounded double-double sums on the GPU) eventually flip one decision."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from oracle import fb_oracle as o  # noqa: E402
from oracle import panoc_oracle as po  # noqa: E402


def ulps(a, b):
    return 0.0 if a == b else abs(a - b) / np.spacing(max(abs(a), abs(b)))


def main():
    out = {}
    for name in ("tiny", "small", "medium"):
        d = np.load(os.path.join(ROOT, "tests", "golden", f"lasso_{name}.npz"))
        A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
        n = A.shape[1]
        it_o = po.PANOCIteration(np.zeros(n), f=o.SquaredDistance(b), A=A, g=o.NormL1(lam))
        it_p = pa.PANOCIteration(np.zeros(n), f=pa.SquaredDistance(b), A=A, g=pa.NormL1(lam))
        tol = 1e-6
        rec = {"first_decision_divergence": None, "iterations_oracle": None, "iterations_gpu": None, "max_rel_iterate_gap_before": 0.0}
        done_o = done_p = None
        for k, (so, sp) in enumerate(zip(it_o, it_p), start=1):
            stop_o = np.max(np.abs(so.res)) / so.gamma <= tol
            stop_p = float(sp.res_norm_inf) / float(sp.gamma) <= tol
            if done_o is None and stop_o:
                done_o = k
            if done_p is None and stop_p:
                done_p = k
            if rec["first_decision_divergence"] is None:
                same = float(sp.gamma) == float(so.gamma) and float(sp.tau) == float(so.tau)
                zg = sp.z.cpu().numpy()
                gap = float(np.max(np.abs(zg - so.z)) / max(1.0, np.max(np.abs(so.z))))
                if same:
                    rec["max_rel_iterate_gap_before"] = max(rec["max_rel_iterate_gap_before"], gap)
                else:
                    to, tp = getattr(so, "line_search_trace", None), getattr(sp, "line_search_trace", None)
                    rec["first_decision_divergence"] = {
                        "iteration": k, "gamma_oracle": float(so.gamma), "gamma_gpu": float(sp.gamma), "tau_oracle": float(so.tau), "tau_gpu": float(sp.tau),
                        "oracle_FBE_x_threshold_FBE_new": to, "gpu_FBE_x_threshold_FBE_new": tp,
                        "ulps_FBE_x": ulps(to[0], tp[0]) if to and tp else None, "ulps_threshold": ulps(to[1], tp[1]) if to and tp else None,
                        "ulps_FBE_new": ulps(to[2], tp[2]) if to and tp else None,
                        "margin_oracle_in_ulps": (to[2] - to[1]) / np.spacing(abs(to[1])) if to else None,
                        "margin_gpu_in_ulps": (tp[2] - tp[1]) / np.spacing(abs(tp[1])) if tp else None,
                        "rel_iterate_gap_at_divergence": gap}
            if (done_o and done_p) or k >= 1000:
                break
        rec["iterations_oracle"], rec["iterations_gpu"] = done_o, done_p
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "panoc_divergence.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
