"""The C port used as the timed CPU baseline computes exactly what the numpy oracle computes (bit for bit)."""
import numpy as np
import pytest

from oracle import fb_oracle as o
from oracle import fb_port


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("prox", ["l1", "box"])
def test_port_iteration_equals_numpy_oracle(T, prox):
    rng = np.random.default_rng(0)
    n = 10_007
    z, zp, grad = (rng.standard_normal(n).astype(T) for _ in range(3))
    gamma, beta = T(0.1), T(0.5)
    if prox == "l1":
        g, kind, p0, p1 = o.NormL1(T(1.0)), fb_port.PROX_L1, 1.0, 0.0
    else:
        g, kind, p0, p1 = o.IndBox(T(-1), T(1)), fb_port.PROX_BOX, -1.0, 1.0
    port = fb_port.FistaPort(z.copy(), zp.copy(), grad, kind, p0, p1)
    zc, zpc = z.copy(), zp.copy()
    for _ in range(3):
        x = zc + beta * (zc - zpc)                      # fast_forward_backward.jl:135
        zpc = zc                                         # :136
        y, zc, res, g_z = o.fb_step_unfused(x, grad, gamma, g)   # :140-142
        rinf, g_z_port = port.step(gamma, beta)
        assert np.array_equal(port.x, x) and np.array_equal(port.y, y)
        assert np.array_equal(port.z, zc) and np.array_equal(port.res, res)
        assert rinf == float(np.max(np.abs(res)))
        assert np.isclose(g_z_port, float(g_z), rtol=1e-5)
    assert fb_port.num_threads() >= 1
    v = np.empty(1000, T)
    fb_port.fill(v, 3)
    assert np.all(np.abs(v) <= 1) and np.std(v) > 0.3
