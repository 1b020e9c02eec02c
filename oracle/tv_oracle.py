"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Anisotropic total-variation denoising by Douglas-Rachford splitting in product-space (consensus) form -- the workload of
BASELINE.json configs[4].  **PARITY UNPINNED**: total variation appears nowhere in the reference (SURVEY.md section 8f, row
f2), so there is no reference arithmetic to follow; the DR iteration itself is the reference's
(src/algorithms/douglas_rachford.jl:54-63, restated in oracle/panoc_oracle.py: DouglasRachfordIteration) and this module only
supplies the two proximable terms it is run with.  Pinned against properties instead (tests/test_oracle_tv.py): each pair
prox is the exact minimiser of its two-variable problem, the fixed point satisfies the TV-denoising optimality conditions
(checked through the dual certificate on small images), and a 1-D signal reproduces the taut-string solution.

    minimize 0.5*||u - b||^2 + lam*( sum_ij |u[i,j+1] - u[i,j]| + sum_ij |u[i+1,j] - u[i,j]| )
    = sum_k f_k(u):  f_0 data term;  f_1/f_2 even/odd horizontal pairs;  f_3/f_4 even/odd vertical pairs
    DR on X = (x_0..x_4):  F(X) = sum_k f_k(x_k) (class TVSplit),  G = indicator{x_0 = ... = x_4} (class Consensus)

Arithmetic (element type T, every operation rounded separately, same order as csrc/tv_kernels.cu):
    data:  (x - b)*w + b,  w = 1/(1 + gamma) rounded once
    pair:  d = a - c; t = gamma*lam;  |d| <= 2t -> (a + c)*0.5 for both, else a - copysign(t, d)
    mean:  ((((r0 + r1) + r2) + r3) + r4)*0.2
"""
from __future__ import annotations

import numpy as np

from .fb_oracle import _R


def _pair(a, c, t):
    T = a.dtype.type
    d = a - c
    return np.where(np.abs(d) <= T(2) * t, (a + c) * T(0.5), a - np.copysign(t, d)).astype(a.dtype)


class TVSplit:
    """F(X) = sum_k f_k(x_k) over the 5 stacked copies of an H x W image (flat vector of 5*H*W, copy-major, row-major image).
    `row0`, `Hglob`: this array holds rows [row0, row0 + H) of a taller image (row-sharded runs): pair parity follows GLOBAL
    row indices; `halo_prev` / `halo_next` are the neighbouring rows (of the copy that needs them) or None at the true border."""

    def __init__(self, b, lam, shape, row0=0, Hglob=None):
        self.H, self.W = shape
        self.b = np.ascontiguousarray(b).reshape(self.H, self.W)
        self.lam = lam
        self.row0 = row0
        self.Hglob = self.H if Hglob is None else Hglob
        self.halo_prev = self.halo_next = None

    def prox(self, X, gamma):
        T = _R(X)
        H, W = self.H, self.W
        x = X.reshape(5, H, W)
        y = x.copy()
        t = T(T(gamma) * T(self.lam))
        w = T(T(1) / T(T(1) + T(gamma)))
        y[0] = ((x[0] - self.b) * w + self.b).astype(T)
        # horizontal pairs: even (0,1),(2,3)...; odd (1,2),(3,4)...
        for k, start in ((1, 0), (2, 1)):
            a, c = x[k][:, start:W - 1:2], x[k][:, start + 1:W:2]
            y[k][:, start:W - 1:2] = _pair(a, c, t)
            y[k][:, start + 1:W:2] = _pair(c, a, t)
        # vertical pairs by GLOBAL row parity
        for k, par in ((3, 0), (4, 1)):
            first = (par - self.row0) % 2           # local index of the first row whose global index has parity `par`
            a, c = x[k][first:H - 1:2], x[k][first + 1:H:2]
            y[k][first:H - 1:2] = _pair(a, c, t)
            y[k][first + 1:H:2] = _pair(c, a, t)
            if first == 1 and self.row0 > 0 and self.halo_prev is not None:       # row 0 pairs with the row above the shard
                y[k][0] = _pair(x[k][0], self.halo_prev, t)
            last_pairs_down = (self.row0 + H - 1) % 2 == par and self.row0 + H < self.Hglob
            if last_pairs_down and self.halo_next is not None:
                y[k][H - 1] = _pair(x[k][H - 1], self.halo_next, t)
        return y.reshape(-1), T(0)

    def objective(self, u):
        u = np.asarray(u, np.float64).reshape(self.H, self.W)
        return (0.5 * np.sum((u - self.b) ** 2) + float(self.lam) * (np.abs(np.diff(u, axis=1)).sum() + np.abs(np.diff(u, axis=0)).sum()))


class Consensus:
    """G = indicator{x_0 = ... = x_{K-1}}: prox = every copy replaced by the mean, summed in copy order."""

    def __init__(self, K=5):
        self.K = K

    def prox(self, R_, gamma):
        T = _R(R_)
        r = R_.reshape(self.K, -1)
        s = r[0].copy()
        for k in range(1, self.K):
            s = (s + r[k]).astype(T)
        z = (s * T(1.0 / self.K)).astype(T)
        return np.tile(z, self.K), T(0)


def tv_denoise_direct_1d(b, lam):
    """Exact 1-D TV denoising (for pinning the splitting on a 1 x W image): solves the dual box-constrained QP
    min 0.5||D' w - b||^2, |w| <= lam by projected gradient to machine precision (tiny sizes only)."""
    b = np.asarray(b, np.float64)
    n = b.size
    D = np.diff(np.eye(n), axis=0)
    w = np.zeros(n - 1)
    for _ in range(200000):
        g = D @ (D.T @ w - b)
        w_new = np.clip(w - 0.25 * g, -lam, lam)
        if np.max(np.abs(w_new - w)) < 1e-15:
            break
        w = w_new
    return b - D.T @ w
