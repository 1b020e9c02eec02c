import math, sys, os, json
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle import fb_oracle as o
from oracle import panoc_oracle as po

def split(a):
    c = 134217729.0 * a
    hi = c - (c - a)
    return hi, a - hi
def exact_dot(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    p = a * b
    ah, al = split(a); bh, bl = split(b)
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return np.float64(math.fsum(p.tolist() + e.tolist()))
def exact_norm2(v):
    return np.float64(math.sqrt(exact_dot(v, v)))

def run(name, exact):
    d = np.load(f'/root/repo/tests/golden/lasso_{name}.npz')
    A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
    n = A.shape[1]
    if exact:
        po.dot, po.norm2 = exact_dot, exact_norm2
    else:
        po.dot, po.norm2 = o.dot, o.norm2
    it = po.PANOCIteration(np.zeros(n), f=o.SquaredDistance(b), A=A, g=o.NormL1(lam))
    out = []
    for k, st in enumerate(it, start=1):
        out.append((float(st.gamma), float(st.tau), st.z.copy(), getattr(st, "line_search_trace", None)))
        if np.max(np.abs(st.res)) / st.gamma <= 1e-6 or k >= 1000:
            break
    return out

res = {}
for name in ("tiny", "small", "medium"):
    a, b_ = run(name, False), run(name, True)
    first = None; gap_before = 0.0
    for k, (u, v) in enumerate(zip(a, b_), start=1):
        gap = float(np.max(np.abs(u[2] - v[2])) / max(1.0, np.max(np.abs(u[2]))))
        if (u[0], u[1]) != (v[0], v[1]):
            first = dict(iteration=k, tau_blas=u[1], tau_exact=v[1], gamma_blas=u[0], gamma_exact=v[0], rel_iterate_gap=gap, trace_blas=u[3], trace_exact=v[3])
            break
        gap_before = max(gap_before, gap)
    res[name] = dict(iterations_blas_dots=len(a), iterations_exact_dots=len(b_), first_decision_divergence=first, max_rel_iterate_gap_before=gap_before)
    print(name, json.dumps(res[name]))
json.dump(res, open('/root/repo/gpurun_out/panoc_oracle_self.json', 'w'), indent=1)
