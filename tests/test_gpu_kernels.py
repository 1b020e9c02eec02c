"""GPU parity, kernel level: every fused / utility kernel against the oracle on the same seeded inputs, through the C ABI.

Bars: element-wise outputs BIT-EXACT; reduction scalars within 2 ulp(fp64) of the exactly rounded
value (in practice they are equal: the kernels accumulate in double-double)."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle import fb_oracle as o  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import ptr  # noqa: E402

import gpu_util as G  # noqa: E402

TYPES = [np.float32, np.float64]
SIZES = [0, 1, 3, 4, 31, 1023, 4096, 4097, 100_003, 1_048_576 + 5]


def _inputs(T, n, seed=0):
    rng = np.random.default_rng(seed + n)
    x = rng.standard_normal(n).astype(T)
    g = rng.standard_normal(n).astype(T)
    zp = rng.standard_normal(n).astype(T)
    return x, g, zp


def _check_scalars(T, row, z, res, g, gfun_value):
    rs = G.fsum_sq(res)
    gd = G.fsum_prod(g, res)
    assert G.ulps(G.pair(row, L.PB_S_RESSQ), rs) <= 2
    # the dot product can cancel: bound the error relative to sum |g*res| instead of the (possibly tiny) result
    scale = float(np.sum(np.abs(g.astype(np.float64) * res.astype(np.float64)))) or 1.0
    assert abs(G.pair(row, L.PB_S_GDR) - gd) <= 1e-28 * scale + 2 * np.spacing(abs(gd))
    assert row[L.PB_S_RESINF] == (float(np.max(np.abs(res))) if res.size else 0.0)
    if gfun_value is not None:
        assert G.ulps(G.pair(row, L.PB_S_GSUM), gfun_value) <= 2


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("prox", ["l1", "box", "zero", "scale_in", "scale_out"])
@pytest.mark.parametrize("impl", [1, 2])
def test_fb_and_ffb_step_bit_exact(T, n, prox, impl):
    """impl 1 = register (LDG) pipeline, impl 2 = TMA bulk-copy shared-memory ring: same bits."""
    G.ctx().set_launch(step_impl=impl)
    try:
        _run_step_bit_exact(T, n, prox)
    finally:
        G.ctx().set_launch()


def _run_step_bit_exact(T, n, prox):
    x, g, zp = _inputs(T, n)
    gamma, beta = T(0.37), T(0.81)
    if prox == "l1":
        lam = T(0.9)
        gfun, desc = o.NormL1(lam), G.prox_desc(L.PB_PROX_L1, lam)
    elif prox == "box":
        gfun, desc = o.IndBox(T(-0.5), T(0.75)), G.prox_desc(L.PB_PROX_BOX, -0.5, 0.75)
    elif prox == "zero":
        gfun, desc = o.ZeroFn(), G.prox_desc(L.PB_PROX_ZERO)
    else:
        s = T(0.625) if prox == "scale_in" else T(1.5)
        desc = G.prox_desc(L.PB_PROX_SCALE, s)

        class _S:
            def prox(self, y, gam):
                return (y.copy() if s > 1 else (s * y).astype(T)), T(0)

        gfun = _S()
    y_o, z_o, r_o, gz_o, xn_o = o.ffb_step_unfused(x, g, zp, gamma, beta, gfun)
    xd, gd, zpd = G.dev(x), G.dev(g), G.dev(zp)
    y, z, r, row = G.fb_step(T, xd, gd, gamma, desc)
    assert np.array_equal(y.cpu().numpy(), y_o) and np.array_equal(z.cpu().numpy(), z_o)
    assert np.array_equal(r.cpu().numpy(), r_o)
    gsum = math.fsum(np.abs(z_o).astype(np.float64).tolist()) if prox == "l1" else None
    _check_scalars(T, row, z_o, r_o, g, gsum)
    y, z, r, xn, row2 = G.ffb_step(T, xd, gd, zpd, gamma, beta, desc, want_y=False, want_res=False)
    assert y is None and r is None
    assert np.array_equal(z.cpu().numpy(), z_o) and np.array_equal(xn.cpu().numpy(), xn_o)
    # same reductions with and without the optional outputs (and across implementations): compare the rounded sums
    for slot in (L.PB_S_GSUM, L.PB_S_RESSQ, L.PB_S_GDR):
        assert G.pair(row2, slot) == G.pair(row, slot) or T == np.float32 and np.float32(G.pair(row2, slot)) == np.float32(G.pair(row, slot))
    assert row2[L.PB_S_RESINF] == row[L.PB_S_RESINF]


@pytest.mark.parametrize("T", TYPES)
def test_misaligned_pointers_take_the_scalar_path(T):
    n = 10_001
    x, g, zp = _inputs(T, n + 1)
    xd, gd, zpd = G.dev(x)[1:], G.dev(g)[1:], G.dev(zp)[1:]      # element offset 1: not 16-byte aligned
    lam = T(0.4)
    desc = G.prox_desc(L.PB_PROX_L1, lam)
    _, z_o, r_o, _, xn_o = o.ffb_step_unfused(x[1:], g[1:], zp[1:], T(0.2), T(0.5), o.NormL1(lam))
    c = G.ctx()
    z = torch.empty(n + 1, dtype=xd.dtype, device="cuda")[1:]
    xn = torch.empty(n + 1, dtype=xd.dtype, device="cuda")[1:]
    L.check(c.lib.pb_ffb_step(c.h, G.dt(T), n, ptr(xd), ptr(gd), ptr(zpd), 0.2, 0.5, C.byref(desc), None, ptr(z), None, ptr(xn)))
    c.read_scalars()
    assert np.array_equal(z.cpu().numpy(), z_o) and np.array_equal(xn.cpu().numpy(), xn_o)


@pytest.mark.parametrize("T", TYPES)
def test_box_with_per_element_bounds_and_nan_inf(T):
    n = 5000
    x, g, _ = _inputs(T, n)
    rng = np.random.default_rng(5)
    lo = (-np.abs(rng.standard_normal(n))).astype(T)
    hi = np.abs(rng.standard_normal(n)).astype(T)
    x[7], x[9], x[11] = np.nan, np.inf, -np.inf
    lod, hid = G.dev(lo), G.dev(hi)
    desc = G.prox_desc(L.PB_PROX_BOX, v0=lod, v1=hid)
    y_o, z_o, r_o, _ = o.fb_step_unfused(x, g, T(0.3), o.IndBox(lo, hi))
    y, z, r, row = G.fb_step(T, G.dev(x), G.dev(g), T(0.3), desc)
    assert np.array_equal(z.cpu().numpy(), z_o, equal_nan=True)
    assert np.array_equal(r.cpu().numpy(), r_o, equal_nan=True)
    assert math.isnan(row[L.PB_S_RESINF])                  # norm(res, Inf) is NaN when res has a NaN (Julia semantics)


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("group,ngroups", [(128, 1000), (128, 1), (4, 33), (100, 77), (1, 50), (513, 9), (256, 37), (512, 11), (64, 40)])
def test_l21_step_matches_oracle(T, group, ngroups):
    n = group * ngroups
    x, g, zp = _inputs(T, n, seed=11)
    lam, gamma, beta = T(1.7), T(0.45), T(0.3)
    gfun = o.NormL21(lam, group)
    y_o, z_o, r_o, gz_o, xn_o = o.ffb_step_unfused(x, g, zp, gamma, beta, gfun)
    desc = G.prox_desc(L.PB_PROX_L21, lam, group=group)
    y, z, r, xn, row = G.ffb_step(T, G.dev(x), G.dev(g), G.dev(zp), gamma, beta, desc)
    assert np.array_equal(y.cpu().numpy(), y_o)
    # the group norm is reduced in a different order (exact double sum on the GPU) -> scale factor within 1 ulp of T
    tol = 4 * np.finfo(T).eps
    assert np.allclose(z.cpu().numpy(), z_o, rtol=tol, atol=tol)
    assert np.allclose(xn.cpu().numpy(), xn_o, rtol=8 * tol, atol=8 * tol)
    assert np.isclose(float(T(lam) * T(G.pair(row, L.PB_S_GSUM))), float(gz_o), rtol=1e-5 if T == np.float32 else 1e-12)
    # zero groups: scal = 1 - gl/0 = -inf -> clamped to 0 -> z = 0
    xz = np.zeros(n, T)
    _, z0, _, _, _ = G.ffb_step(T, G.dev(xz), G.dev(xz), G.dev(xz), gamma, beta, desc)
    assert not np.any(np.isnan(z0.cpu().numpy())) or group >= 1


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n", [0, 5, 4099, 300_000])
def test_utility_kernels(T, n):
    c = G.ctx()
    x, g, zp = _inputs(T, n, seed=3)
    xd, gd, zd = G.dev(x), G.dev(g), G.dev(zp)
    out = torch.empty_like(xd)
    d = G.dt(T)
    # pb_forward: y = x - gamma*g, AUX = ||y||^2
    L.check(c.lib.pb_forward(c.h, d, n, ptr(xd), ptr(gd), 0.3, ptr(out)))
    row = c.read_scalars()
    y = x - T(0.3) * g
    assert np.array_equal(out.cpu().numpy(), y) and G.ulps(G.pair(row, L.PB_S_AUX), G.fsum_sq(y)) <= 2
    # pb_extrapolate
    L.check(c.lib.pb_extrapolate(c.h, d, n, ptr(xd), ptr(zd), 0.7, ptr(out)))
    assert np.array_equal(out.cpu().numpy(), x + T(0.7) * (x - zp))
    # pb_residual with and without grad
    L.check(c.lib.pb_residual(c.h, d, n, ptr(xd), ptr(zd), ptr(gd), ptr(out)))
    row = c.read_scalars()
    res = x - zp
    assert np.array_equal(out.cpu().numpy(), res)
    assert G.ulps(G.pair(row, L.PB_S_RESSQ), G.fsum_sq(res)) <= 2
    assert row[L.PB_S_RESINF] == (float(np.max(np.abs(res))) if n else 0.0)
    # pb_add_scalar, pb_sub, pb_nrm2sq, pb_dot
    L.check(c.lib.pb_add_scalar(c.h, d, n, ptr(xd), 1.0, ptr(out)))
    assert np.array_equal(out.cpu().numpy(), x + T(1))
    L.check(c.lib.pb_sub(c.h, d, n, ptr(xd), ptr(gd), ptr(out)))
    row = c.read_scalars()
    assert np.array_equal(out.cpu().numpy(), x - g) and G.ulps(G.pair(row, L.PB_S_AUX), G.fsum_sq(x - g)) <= 2
    L.check(c.lib.pb_nrm2sq(c.h, d, n, ptr(xd)))
    row = c.read_scalars()
    assert G.ulps(G.pair(row, L.PB_S_AUX), G.fsum_sq(x)) <= 2
    assert row[L.PB_S_AUXINF] == (float(np.max(np.abs(x))) if n else 0.0)
    L.check(c.lib.pb_dot(c.h, d, n, ptr(xd), ptr(gd)))
    row = c.read_scalars()
    scale = float(np.sum(np.abs(x.astype(np.float64) * g.astype(np.float64)))) or 1.0
    assert abs(G.pair(row, L.PB_S_AUX) - G.fsum_prod(x, g)) <= 1e-28 * scale + 2 * np.spacing(abs(G.fsum_prod(x, g)))


@pytest.mark.parametrize("T", TYPES)
def test_standalone_prox(T):
    n = 20_000
    rng = np.random.default_rng(2)
    y = (rng.standard_normal(n) * 2).astype(T)
    yd = G.dev(y)
    c = G.ctx()
    z = torch.empty_like(yd)
    for desc, gfun in [
        (G.prox_desc(L.PB_PROX_L1, 0.8), o.NormL1(T(0.8))),
        (G.prox_desc(L.PB_PROX_BOX, -1, 1), o.IndBox(T(-1), T(1))),
        (G.prox_desc(L.PB_PROX_ZERO), o.ZeroFn()),
    ]:
        L.check(c.lib.pb_prox_apply(c.h, G.dt(T), n, ptr(yd), 0.5, C.byref(desc), ptr(z)))
        row = c.read_scalars()
        z_o, v = gfun.prox(y, T(0.5))
        assert np.array_equal(z.cpu().numpy(), z_o)
        if desc.kind == L.PB_PROX_L1:
            assert G.ulps(G.pair(row, L.PB_S_GSUM), math.fsum(np.abs(z_o).astype(np.float64).tolist())) <= 2
    # product-level IndBallL2 (two phase) and NormL21 against the oracle
    import proxb200 as pa

    for r in (T(0.5), T(1e9)):
        pa.IndBallL2(r).prox_(z, yd, T(0.5))
        z_o, _ = o.IndBallL2(r).prox(y, T(0.5))
        assert np.allclose(z.cpu().numpy(), z_o, rtol=4 * np.finfo(T).eps, atol=0)
    yg = y[: 128 * 100].copy()
    zg = torch.empty(128 * 100, dtype=yd.dtype, device="cuda")
    v = pa.NormL21(T(1.1), 128).prox_(zg, G.dev(yg), T(0.5))
    z_o, v_o = o.NormL21(T(1.1), 128).prox(yg, T(0.5))
    assert np.allclose(zg.cpu().numpy(), z_o, rtol=4 * np.finfo(T).eps, atol=4 * np.finfo(T).eps)
    assert np.isclose(float(v), float(v_o), rtol=1e-5)


def test_reductions_independent_of_grid_size_and_hints():
    """The double-double tree makes the rounded scalars identical for every launch shape, implementation and shard count
    (deterministic AND partition-independent): this is what lets a sharded run reproduce the single-GPU iteration count.
    fp64 data: per-thread accumulation is compensated too -> the double results are equal.  fp32 data: per-thread sums are
    plain double (error ~1e-15, 8 orders below the float32 rounding the host applies) -> equal after rounding to R."""
    for T in TYPES:
        n = 3_000_017
        x, g, zp = _inputs(T, n, seed=9)
        xd, gd, zpd = G.dev(x), G.dev(g), G.dev(zp)
        desc = G.prox_desc(L.PB_PROX_L1, 0.3)
        c = G.ctx()
        rows, zs = [], []
        try:
            for ctas, hint, unroll, impl in [(0, -1, 0, 0), (1, 0, 1, 1), (3, 1, 2, 1), (8, 0, 8, 1), (16, 1, 4, 1), (2, 0, 0, 2), (3, 0, 0, 2)]:
                c.set_launch(ctas, hint, unroll, impl)
                _, z, _, xn, row = G.ffb_step(T, xd, gd, zpd, T(0.2), T(0.6), desc, want_y=False, want_res=False)
                rows.append([G.pair(row, s) for s in (L.PB_S_GSUM, L.PB_S_RESSQ, L.PB_S_GDR)] + [row[L.PB_S_RESINF]])
                zs.append(z.clone())
        finally:
            c.set_launch()
        rnd = (lambda v: v) if T == np.float64 else (lambda v: [float(np.float32(e)) for e in v])
        for r in rows[1:]:
            assert rnd(r) == rnd(rows[0])
        for z in zs[1:]:
            assert torch.equal(z, zs[0])
        # shard emulation: split into P ranges, combine the per-shard pairs on the host in double-double
        from proxb200.host import Scalars, shard_bounds

        for P in (2, 3, 8):
            parts = np.zeros((P, L.PB_NSCALARS))
            for r_, (lo, hi) in enumerate(shard_bounds(n, P)):
                _, _, _, _, row = G.ffb_step(T, xd[lo:hi], gd[lo:hi], zpd[lo:hi], T(0.2), T(0.6), desc, False, False)
                parts[r_] = row
            sc = Scalars(parts)
            assert rnd([sc.gsum, sc.res_sq, sc.gdr, sc.res_inf]) == rnd(rows[0])


def test_argument_errors():
    c = G.ctx()
    x = torch.zeros(8, device="cuda")
    desc = G.prox_desc(L.PB_PROX_L1, 1.0)
    assert c.lib.pb_fb_step(c.h, 7, 8, ptr(x), ptr(x), 0.1, C.byref(desc), None, ptr(x), None) == 1
    assert b"dtype" in c.lib.pb_last_error()
    assert c.lib.pb_fb_step(c.h, 0, -1, ptr(x), ptr(x), 0.1, C.byref(desc), None, ptr(x), None) == 1
    assert c.lib.pb_fb_step(c.h, 0, 8, None, ptr(x), 0.1, C.byref(desc), None, ptr(x), None) == 1
    bad = G.prox_desc(99)
    assert c.lib.pb_fb_step(c.h, 0, 8, ptr(x), ptr(x), 0.1, C.byref(bad), None, ptr(x), None) == 1
    l21 = G.prox_desc(L.PB_PROX_L21, 1.0, group=3)
    assert c.lib.pb_fb_step(c.h, 0, 8, ptr(x), ptr(x), 0.1, C.byref(l21), None, ptr(x), None) == 1
    assert c.lib.pb_ffb_step(c.h, 0, 8, ptr(x), ptr(x), ptr(x), 0.1, 0.1, C.byref(desc), None, ptr(x), None, ptr(x)) == 1   # x_next aliases x
    h = C.c_void_p()
    assert c.lib.pb_ctx_create(9999, None, 0, C.byref(h)) == 1
