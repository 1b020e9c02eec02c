"""Pins the (parity-unpinned, see oracle/tv_oracle.py) TV splitting through its mathematical properties.  CPU only."""
import numpy as np
import pytest

from oracle import panoc_oracle as po
from oracle import tv_oracle as tv


def test_pair_prox_is_the_minimiser():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, c, t = rng.standard_normal(), rng.standard_normal(), abs(rng.standard_normal()) * 0.5
        pa_, pc_ = tv._pair(np.array([a]), np.array([c]), np.float64(t))[0], tv._pair(np.array([c]), np.array([a]), np.float64(t))[0]
        obj = lambda u, v: t * abs(u - v) + 0.5 * ((u - a) ** 2 + (v - c) ** 2)   # noqa: E731
        best = obj(pa_, pc_)
        for du, dv in rng.standard_normal((50, 2)) * 1e-3:
            assert obj(pa_ + du, pc_ + dv) >= best - 1e-15


def test_dr_tv_1d_matches_the_exact_solution():
    rng = np.random.default_rng(1)
    W = 24
    b = np.cumsum(rng.standard_normal(W) * (rng.random(W) < 0.3)) + 0.1 * rng.standard_normal(W)
    lam = 0.4
    f, g = tv.TVSplit(b, lam, (1, W)), tv.Consensus(5)
    it = po.DouglasRachfordIteration(np.tile(b, 5), f=f, g=g, gamma=1.0)
    for k, st in enumerate(it):
        if k == 3000:
            break
    u = st.z[:W]
    want = tv.tv_denoise_direct_1d(b, lam)
    assert np.max(np.abs(u - want)) <= 1e-6
    assert np.max(np.abs(st.y.reshape(5, W) - u)) <= 1e-6                # consensus reached: every copy equals u


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_dr_tv_2d_decreases_the_objective_and_shards_consistently(T):
    rng = np.random.default_rng(2)
    H, W = 12, 10
    img = np.zeros((H, W))
    img[3:9, 2:7] = 1.0
    b = (img + 0.2 * rng.standard_normal((H, W))).astype(T)
    lam, gamma = T(0.3), T(1.0)
    f, g = tv.TVSplit(b, lam, (H, W)), tv.Consensus(5)
    x0 = np.tile(b.reshape(-1), 5)
    it = iter(po.DouglasRachfordIteration(x0, f=f, g=g, gamma=gamma))
    objs = []
    for k in range(400):
        st = next(it)
        if k % 100 == 99:
            objs.append(f.objective(st.z[: H * W]))
    assert objs[-1] <= objs[0] + 1e-6 and objs[-1] < f.objective(b)
    # a two-shard run with exchanged halo rows reproduces the unsharded iterates bit for bit
    Hs = 6
    shards = [tv.TVSplit(b[:Hs], lam, (Hs, W), 0, H), tv.TVSplit(b[Hs:], lam, (H - Hs, W), Hs, H)]
    X = [np.tile(b[:Hs].reshape(-1), 5), np.tile(b[Hs:].reshape(-1), 5)]
    Xf = x0.copy()
    for _ in range(5):
        top, bot = X[0].reshape(5, Hs, W), X[1].reshape(5, H - Hs, W)
        kc = 3 if (Hs - 1) % 2 == 0 else 4                              # the copy whose pair straddles the boundary
        shards[0].halo_next, shards[1].halo_prev = bot[kc][0].copy(), top[kc][Hs - 1].copy()
        new = []
        for s, xs in zip(shards, X):
            y, _ = s.prox(xs, gamma)
            r = (2 * y - xs).astype(T)
            z, _ = g.prox(r, gamma)
            new.append((xs - (y - z)).astype(T))
        X = new
        y, _ = f.prox(Xf, gamma)
        r = (2 * y - Xf).astype(T)
        z, _ = g.prox(r, gamma)
        Xf = (Xf - (y - z)).astype(T)
        full = Xf.reshape(5, H, W)
        assert np.array_equal(X[0].reshape(5, Hs, W), full[:, :Hs]) and np.array_equal(X[1].reshape(5, H - Hs, W), full[:, Hs:])
