"""The JLD2/HDF5 fixture reader (row f3 of SURVEY.md section 8f) against the committed golden arrays.  Needs the reference tree
(present in the build container, absent on the GPU box): skipped there."""
import os

import numpy as np
import pytest

from proxb200 import jld2

REF = os.environ.get("PROX_REFERENCE", "/root/reference")


@pytest.mark.parametrize("name", ["tiny", "small", "medium"])
def test_reader_matches_golden(golden, name):
    path = os.path.join(REF, "benchmark", "data", f"lasso_{name}.jld2")
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    A, b, lam, xstar, ystar = jld2.load_lasso_fixture(path)
    g = golden("lasso_" + name)
    assert A.flags.f_contiguous and A.dtype == np.float64
    assert np.array_equal(A, g["A"]) and np.array_equal(b, g["b"]) and lam == float(g["lam"])
    assert np.array_equal(xstar, g["xstar"]) and np.array_equal(ystar, g["ystar"])
    assert np.allclose(ystar, b - A @ xstar, atol=1e-12)         # the relation the survey found in the files


def test_reader_rejects_other_files(tmp_path):
    p = tmp_path / "x.jld2"
    p.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(ValueError):
        jld2.read_jld2(str(p))
