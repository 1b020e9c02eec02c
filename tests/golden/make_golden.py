#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the reference tree.

Runs ONLY in the build container (needs /root/reference); the GPU box never has the reference,
so everything the tests need is written to small .npz files that are committed next to this script.

Sources (all paths relative to /root/reference):
  * benchmark/data/lasso_{tiny,small,medium}.jld2  -- HDF5-in-JLD2 fixtures (A, b, lambda, xstar, ystar).
    No HDF5 reader exists in this image, so the little-endian f64 payloads are read at the byte offsets
    established in SURVEY.md section 8c; each file's sha256 prefix is asserted so a different file cannot
    be mis-read silently.
  * test/problems/test_lasso_small.jl:17-23,42  -- literal 4x5 A, b and x_star.
  * test/problems/test_lasso_small_strongly_convex.jl:11-44,53 -- literal w, B, x_star; A = Q D Q',
    b = A x* + lam inv(A') sign(x*), x0 = A \\ b are re-derived with LAPACK exactly as the test does.
  * test/accel/test_lbfgs.jl:6-101 -- literal Q, q, xs and the five reference L-BFGS directions.
  * test/problems/test_sparse_logistic_small.jl:34 -- literal x_star.
  * test/problems/test_elasticnet.jl:27 -- literal x_star.
  * test/problems/test_linear_programs.jl:44-96 -- literal LP data (x_star, s_star, y_star, A).
The literals are parsed out of the .jl files with regular expressions (nothing is copied by hand).
"""
import hashlib
import os
import re
import sys

import numpy as np

REF = os.environ.get("PROX_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

# name: (rows, cols, off_A, off_ystar, off_b, off_lambda, off_xstar, sha256[:16])
JLD2 = {
    "tiny": (5, 10, 630, 1095, 1200, 1289, 1362, "b5231e9c4c793d64"),
    "small": (50, 100, 647, 40709, 41175, 41624, 41698, "9dd5ab8fc9720ac1"),
    "medium": (500, 1000, 647, 4000709, 4004775, 4008824, 4008898, "253e738baaa339cd"),
}


def read_jld2(name):
    m, n, o_a, o_y, o_b, o_l, o_x, sha = JLD2[name]
    path = os.path.join(REF, "benchmark", "data", f"lasso_{name}.jld2")
    buf = open(path, "rb").read()
    got = hashlib.sha256(buf).hexdigest()[:16]
    if got != sha:
        raise SystemExit(f"{path}: sha256 {got} != expected {sha}; offsets would be wrong")
    a = np.frombuffer(buf, "<f8", m * n, o_a).reshape(n, m).T  # Julia column-major
    # cross-check: the package's own HDF5/JLD2 parser (proximalalgorithms.jl_b200/jld2.py) finds the same datasets
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from proxb200 import jld2 as _jld2

    parsed = _jld2.read_jld2(path)
    assert np.array_equal(parsed["A"], a) and parsed["A"].shape == (m, n) and int(parsed["lambda"]) == int(np.frombuffer(buf, "<i8", 1, o_l)[0])
    assert np.array_equal(parsed["b"], np.frombuffer(buf, "<f8", m, o_b)) and np.array_equal(parsed["xstar"], np.frombuffer(buf, "<f8", n, o_x))
    return dict(
        A=np.asfortranarray(a),
        b=np.frombuffer(buf, "<f8", m, o_b).copy(),
        ystar=np.frombuffer(buf, "<f8", m, o_y).copy(),
        xstar=np.frombuffer(buf, "<f8", n, o_x).copy(),
        lam=np.float64(np.frombuffer(buf, "<i8", 1, o_l)[0]),
    )


_NUM = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?"


def _block_after(text, marker):
    """Return the text between the '[' that follows `marker` and its closing ']'."""
    i = text.index(marker)
    i = text.index("[", i)
    j = text.index("]", i)
    return text[i + 1 : j]


def _matrix(block):
    rows = [r for r in block.strip().splitlines() if r.strip()]
    return np.array([[float(t) for t in re.findall(_NUM, r)] for r in rows], dtype=np.float64)


def _vector(block):
    return np.array([float(t) for t in re.findall(_NUM, block)], dtype=np.float64)


def unit_lasso_4x5():
    src = open(os.path.join(REF, "test", "problems", "test_lasso_small.jl")).read()
    a = _matrix(_block_after(src, "A = T["))
    b = _vector(_block_after(src, "b = T["))
    xs = _vector(_block_after(src, "x_star = T["))
    assert a.shape == (4, 5) and b.shape == (4,) and xs.shape == (5,)
    return dict(A=np.asfortranarray(a), b=b, xstar=xs)


def unit_lasso_sc_5x5():
    src = open(os.path.join(REF, "test", "problems", "test_lasso_small_strongly_convex.jl")).read()
    xs = _vector(_block_after(src, "x_star = T["))
    w = _vector(_block_after(src, "w = T["))
    bmat = _matrix(_block_after(src, "B = T["))
    assert xs.shape == (5,) and w.shape == (5,) and bmat.shape == (5, 5)
    mf, lf = 1.0, 10.0
    lam = (mf + lf) / 2
    d = np.sqrt(mf) + (np.sqrt(lf) - np.sqrt(mf)) * w
    d[0] = np.sqrt(mf)
    d[-1] = np.sqrt(lf)
    q, _ = np.linalg.qr(bmat)
    a = q @ np.diag(d) @ q.T
    b = a @ xs + lam * np.linalg.solve(a.T, np.sign(xs))
    x0 = np.linalg.solve(a, b)
    return dict(A=np.asfortranarray(a), b=b, xstar=xs, x0=x0, lam=np.float64(lam), mf=np.float64(mf), Lf=np.float64(lf))


def lbfgs_known_answers():
    """test/accel/test_lbfgs.jl:6-101: Q (10x10), q, the five points xs and the five reference directions (memory 3)."""
    src = open(os.path.join(REF, "test", "accel", "test_lbfgs.jl")).read()
    qm = _matrix(_block_after(src, "Q = T["))
    qv = _vector(_block_after(src, "q = T["))
    i = src.index("xs = [")
    j = src.index("dirs_ref = [")
    k = src.index("@testset \"Arrays\"")
    xs = np.array([_vector(b) for b in re.findall(r"T\[(.*?)\]", src[i:j], flags=re.S)])
    dirs = np.array([_vector(b) for b in re.findall(r"T\[(.*?)\]", src[j:k], flags=re.S)])
    assert qm.shape == (10, 10) and qv.shape == (10,) and xs.shape == (5, 10) and dirs.shape == (5, 10)
    return dict(Q=qm, q=qv, xs=xs, dirs_ref=dirs)


def unit_sparse_logistic():
    """test/problems/test_sparse_logistic_small.jl:34 literal x_star (A, b are those of the 4x5 Lasso; lam = 0.1)."""
    src = open(os.path.join(REF, "test", "problems", "test_sparse_logistic_small.jl")).read()
    xs = _vector(_block_after(src, "x_star = T["))
    assert xs.shape == (5,)
    return dict(xstar=xs, lam=np.float64(0.1))


def unit_elasticnet():
    """test/problems/test_elasticnet.jl:27 literal x_star (A, b are those of the 4x5 Lasso; ElasticNet(1, 1))."""
    src = open(os.path.join(REF, "test", "problems", "test_elasticnet.jl")).read()
    xs = _vector(_block_after(src, "x_star = T["))
    assert xs.shape == (5,)
    return dict(xstar=xs)


def unit_linear_program():
    """test/problems/test_linear_programs.jl:44-96: literal x_star, s_star, y_star, A (8 x 10); b = A x*, c = A' y* + s*."""
    src = open(os.path.join(REF, "test", "problems", "test_linear_programs.jl")).read()
    xs = _vector(_block_after(src, "x_star = T["))
    ss = _vector(_block_after(src, "s_star = T["))
    ys = _vector(_block_after(src, "y_star = T["))
    a = _matrix(_block_after(src, "A = T["))
    assert xs.shape == (10,) and ss.shape == (10,) and ys.shape == (8,) and a.shape == (8, 10)
    return dict(A=np.asfortranarray(a), xstar=xs, sstar=ss, ystar=ys, b=a @ xs, c=a.T @ ys + ss)


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: golden fixtures can only be regenerated in the build container")
    for name in JLD2:
        d = read_jld2(name)
        np.savez_compressed(os.path.join(OUT, f"lasso_{name}.npz"), **d)
        r = d["A"] @ d["xstar"] - d["b"]
        print(f"lasso_{name}: A{d['A'].shape} lam={d['lam']} obj(x*)={0.5 * r @ r + d['lam'] * np.abs(d['xstar']).sum():.16g}")
    np.savez_compressed(os.path.join(OUT, "unit_lasso_4x5.npz"), **unit_lasso_4x5())
    np.savez_compressed(os.path.join(OUT, "unit_lasso_sc_5x5.npz"), **unit_lasso_sc_5x5())
    np.savez_compressed(os.path.join(OUT, "lbfgs_known_answers.npz"), **lbfgs_known_answers())
    np.savez_compressed(os.path.join(OUT, "unit_sparse_logistic.npz"), **unit_sparse_logistic())
    np.savez_compressed(os.path.join(OUT, "unit_elasticnet.npz"), **unit_elasticnet())
    np.savez_compressed(os.path.join(OUT, "unit_linear_program.npz"), **unit_linear_program())
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
