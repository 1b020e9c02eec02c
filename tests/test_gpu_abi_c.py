"""The C ABI driven from plain C (tests/c/abi_smoke.c, compiled here with gcc against include/proxb200.h and the in-tree
libproxb200.so): fused step checked against the same arithmetic in C, and a whole FastForwardBackward solve of the reference's
lasso_small fixture (benchmark/benchmarks.jl:55-61) -- iteration count equal to the oracle's, objective equal to the fixture's."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
LIBDIR = os.path.join(ROOT, "proximalalgorithms.jl_b200", "lib")


def _compile(tmp_path, lib_built):
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-ffp-contract=off", SRC, "-I", os.path.join(ROOT, "include"), "-L", LIBDIR, "-lproxb200",
           "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_valid_c_and_links(tmp_path, lib_built):
    """CPU half: proxb200.h compiles as C99 with -Wall -Werror and every symbol the program uses resolves against the library."""
    _compile(tmp_path, lib_built)


@pytest.mark.gpu
def test_c_host_runs_the_fused_step_and_a_whole_solve(tmp_path, lib_built):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = _compile(tmp_path, lib_built)
    d = load_golden("lasso_small")
    A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
    m, n = A.shape
    path = tmp_path / "problem.bin"
    with open(path, "wb") as fh:
        fh.write(struct.pack("<qqd", m, n, lam))
        fh.write(A.tobytes(order="F"))
        fh.write(b.tobytes())
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    mt = re.search(r"iterations=(\d+) persistent_ctas=(\d+) objective=(\S+)", r.stdout)
    assert mt, r.stdout
    assert int(mt.group(1)) == 788                      # the oracle's count on lasso_small (tests/test_gpu_solvers.py)
    obj_star = 0.5 * np.sum((d["A"] @ d["xstar"] - b) ** 2) + lam * np.sum(np.abs(d["xstar"]))
    assert abs(float(mt.group(3)) - obj_star) <= 1e-6 * obj_star
