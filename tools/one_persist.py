#!/usr/bin/env python
"""One persistent-kernel solve of a reference fixture, for `ncu -k regex:k_persist_solve` (profiles/): python tools/one_persist.py small fb"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "small"
alg = sys.argv[2] if len(sys.argv) > 2 else "fb"
d = np.load(os.path.join(ROOT, "tests", "golden", f"lasso_{name}.npz"))
A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
f, g = pa.LeastSquares(A, b), pa.NormL1(lam)
mk = pa.ForwardBackward if alg == "fb" else pa.FastForwardBackward
for _ in range(2):
    s = mk(tol=1e-6, driver="native")
    z, it = s(x0=np.zeros(A.shape[1]), f=f, g=g)
print(name, alg, it, s.last_persistent_ctas)
