"""TEST INFRASTRUCTURE ONLY: a numpy stand-in for libproxb200's C ABI so that the HOST logic of the solvers (buffer renames,
scalar arithmetic in R, line-search control flow) can be exercised on a machine without a GPU.

It is installed by the `emu` fixture (tests/test_host_emulated.py) by monkey-patching `Context.get`; nothing in the product
imports it, and the product itself still refuses to run without CUDA.  Each function follows the contract written in
include/proxb200.h: same argument order, same scalar-block slots, products and sums rounded separately in the element type,
reductions exactly rounded (math.fsum) like the library's double-double sums.
"""
import ctypes as C
import math

import numpy as np
import torch

from proxb200 import _lib as L


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, C.c_void_p):
        return p.value or 0
    if isinstance(p, int):
        return p
    return C.cast(p, C.c_void_p).value or 0


def _vec(p, n, dt):
    a = _addr(p)
    T = np.float32 if dt == L.PB_F32 else np.float64
    if n == 0:
        return np.zeros(0, T)            # empty tensors have a null data pointer
    if a == 0:
        return None
    ct = C.c_float if T is np.float32 else C.c_double
    return np.ctypeslib.as_array((ct * int(n)).from_address(a))


def _fsum(v):
    return math.fsum(np.asarray(v, dtype=np.float64).tolist())


def _fsum_prod(a, b):
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.dtype == np.float32:
        return math.fsum((a64 * b64).tolist())
    # float64: split the products error-free (Dekker) so that the sum is still exactly rounded
    p = a64 * b64
    err = np.array([math.fma(x, y, -q) for x, y, q in zip(a64.tolist(), b64.tolist(), p.tolist())]) if hasattr(math, "fma") else 0 * p
    return math.fsum(p.tolist() + np.asarray(err).tolist())


class _Lbfgs:
    def __init__(self, dt, n, mem):
        self.dt, self.n, self.mem = dt, n, mem
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.T = T
        self.currmem = self.curridx = 0
        self.slot_of = [-1] * (mem + 1)
        self.spare = 0
        self.pending = False
        self.s = np.zeros((mem + 1, max(n, 1)), T)
        self.y = np.zeros((mem + 1, max(n, 1)), T)
        self.ys = [T(0)] * (mem + 1)
        self.H = T(1)
        self.alpha = [T(0)] * (mem + 1)


class EmuLib:
    def __init__(self):
        self.scal = np.zeros(L.PB_NSCALARS)
        self.launches = 0
        self.err = b""
        self._lb = {}
        self._next = 1

    # ---- plumbing ------------------------------------------------------------------------------------------------
    def pb_last_error(self):
        return self.err

    def pb_ctx_launch_count(self, h):
        return self.launches

    def pb_read_scalars(self, h, out):
        for k in range(L.PB_NSCALARS):
            out[k] = self.scal[k]
        return 0

    def _set(self, slot, v):
        self.scal[slot], self.scal[slot + 1] = float(v), 0.0

    def pb_copy(self, h, dst, src, nbytes):
        C.memmove(_addr(dst), _addr(src), nbytes)
        return 0

    # ---- prox ----------------------------------------------------------------------------------------------------
    def _prox(self, T, g, y, gamma, set_gsum=True):
        g = g._obj if hasattr(g, "_obj") else g
        kind = g.kind
        n = y.shape[0]
        if kind == L.PB_PROX_ZERO:
            return y.copy()
        if kind == L.PB_PROX_L1:
            gl = T(T(gamma) * T(g.p0))
            z = (y + np.where(y <= -gl, gl, np.where(y >= gl, -gl, -y)).astype(T)).astype(T)
            if set_gsum:
                self._set(L.PB_S_GSUM, _fsum(np.abs(z)))
            return z
        if kind == L.PB_PROX_BOX:
            lo = _vec(g.v0, n, L.PB_F32 if T is np.float32 else L.PB_F64) if g.v0 else T(g.p0)
            hi = _vec(g.v1, n, L.PB_F32 if T is np.float32 else L.PB_F64) if g.v1 else T(g.p1)
            return np.where(y < lo, lo, np.where(y > hi, hi, y)).astype(T)
        if kind == L.PB_PROX_SCALE:
            s = T(g.p0)
            return y.copy() if s > 1 else (s * y).astype(T)
        if kind == L.PB_PROX_BALL:          # one-GPU IndBallL2: scale factor from ||y||^2 (the library forms it on the device)
            ny = T(math.sqrt(_fsum_prod(y, y)))
            with np.errstate(divide="ignore"):
                sc = T(T(g.p0) / ny)
            return y.copy() if sc > 1 else (sc * y).astype(T)
        if kind == L.PB_PROX_L21:
            gl = T(T(gamma) * T(g.p0))
            yg = y.reshape(-1, g.group)
            ns = np.array([math.sqrt(_fsum_prod(r, r)) for r in yg]).astype(T)
            with np.errstate(divide="ignore", invalid="ignore"):
                scal = (T(1) - gl / ns).astype(T)
            scal = np.where(scal <= 0, T(0), scal).astype(T)
            if set_gsum:
                self._set(L.PB_S_GSUM, _fsum((scal * ns).astype(T)))
            return (scal[:, None] * yg).astype(T).reshape(-1)
        if kind == L.PB_PROX_SQRL2:
            den = T(T(1) + T(T(gamma) * T(g.p0)))
            b = _vec(g.v0, n, L.PB_F32 if T is np.float32 else L.PB_F64) if g.v0 else None
            w = ((y - b) / den).astype(T) if b is not None else (y / den).astype(T)
            if set_gsum:
                self._set(L.PB_S_GSUM, _fsum_prod(w, w))
            return (w + b).astype(T) if b is not None else w
        raise ValueError(kind)

    def pb_prox_apply(self, h, dt, n, y, gamma, g, z):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        _vec(z, n, dt)[...] = self._prox(T, g, _vec(y, n, dt), gamma)
        return 0

    # ---- K1 / K2 -------------------------------------------------------------------------------------------------
    def pb_fb_step(self, h, dt, n, x, grad, gamma, g, y, z, res):
        return self.pb_ffb_step(h, dt, n, x, grad, None, gamma, 0.0, g, y, z, res, None)

    def pb_ffb_step(self, h, dt, n, x, grad, z_prev, gamma, beta, g, y, z, res, x_next):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        xv, gv = _vec(x, n, dt), _vec(grad, n, dt)
        yv = (xv - (T(gamma) * gv).astype(T)).astype(T)
        self._set(L.PB_S_GSUM, 0.0)
        zv = self._prox(T, g, yv, gamma)
        rv = (xv - zv).astype(T)
        if _addr(x_next):
            zp = _vec(z_prev, n, dt)
            _vec(x_next, n, dt)[...] = (zv + (T(beta) * (zv - zp).astype(T)).astype(T)).astype(T)
        _vec(z, n, dt)[...] = zv
        if _addr(y):
            _vec(y, n, dt)[...] = yv
        if _addr(res):
            _vec(res, n, dt)[...] = rv
        self._set(L.PB_S_RESSQ, _fsum_prod(rv, rv))
        self._set(L.PB_S_GDR, _fsum_prod(gv, rv))
        self.scal[L.PB_S_RESINF] = float(np.max(np.abs(rv))) if n else 0.0
        return 0

    # ---- K6 ------------------------------------------------------------------------------------------------------
    def pb_forward(self, h, dt, n, x, grad, gamma, y):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        yv = (_vec(x, n, dt) - (T(gamma) * _vec(grad, n, dt)).astype(T)).astype(T)
        _vec(y, n, dt)[...] = yv
        self._set(L.PB_S_AUX, _fsum_prod(yv, yv))
        return 0

    def pb_extrapolate(self, h, dt, n, z, z_prev, beta, x):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        zv, zp = _vec(z, n, dt), _vec(z_prev, n, dt)
        _vec(x, n, dt)[...] = (zv + (T(beta) * (zv - zp).astype(T)).astype(T)).astype(T)
        return 0

    def pb_residual(self, h, dt, n, x, z, grad, res):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        rv = (_vec(x, n, dt) - _vec(z, n, dt)).astype(T)
        if _addr(res):
            _vec(res, n, dt)[...] = rv
        self._set(L.PB_S_RESSQ, _fsum_prod(rv, rv))
        if _addr(grad):
            self._set(L.PB_S_GDR, _fsum_prod(_vec(grad, n, dt), rv))
        self.scal[L.PB_S_RESINF] = float(np.max(np.abs(rv))) if n else 0.0
        return 0

    def pb_add_scalar(self, h, dt, n, x, c, out):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        _vec(out, n, dt)[...] = (_vec(x, n, dt) + T(c)).astype(T)
        return 0

    def pb_sub(self, h, dt, n, a, b, out):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        o = (_vec(a, n, dt) - _vec(b, n, dt)).astype(T)
        _vec(out, n, dt)[...] = o
        self._set(L.PB_S_AUX, _fsum_prod(o, o))
        return 0

    def pb_nrm2sq(self, h, dt, n, v):
        self.launches += 1
        vv = _vec(v, n, dt)
        self._set(L.PB_S_AUX, _fsum_prod(vv, vv))
        self.scal[L.PB_S_AUXINF] = float(np.max(np.abs(vv))) if n else 0.0
        return 0

    def pb_dot(self, h, dt, n, a, b):
        self.launches += 1
        self._set(L.PB_S_AUX, _fsum_prod(_vec(a, n, dt), _vec(b, n, dt)))
        return 0

    def pb_lincomb2(self, h, dt, n, a, x, b, y, out):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        o = ((T(a) * _vec(x, n, dt)).astype(T) + (T(b) * _vec(y, n, dt)).astype(T)).astype(T)
        _vec(out, n, dt)[...] = o
        return 0

    def pb_scale(self, h, dt, n, s, x, out):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        _vec(out, n, dt)[...] = (T(s) * _vec(x, n, dt)).astype(T)
        return 0

    # ---- K4 ------------------------------------------------------------------------------------------------------
    def pb_lsq_dense_residual(self, h, dt, m, n, A, lda, x, b, r):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        Am = _vec(A, m * n, dt).reshape(n, m).T
        rv = (Am.astype(np.float64) @ _vec(x, n, dt).astype(np.float64))
        if _addr(b):
            rv = rv - _vec(b, m, dt)
        rv = rv.astype(T)
        _vec(r, m, dt)[...] = rv
        self._set(L.PB_S_AUX, _fsum_prod(rv, rv))
        return 0

    def pb_lsq_dense_gradient(self, h, dt, m, n, A, lda, r, grad):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        Am = _vec(A, m * n, dt).reshape(n, m).T
        _vec(grad, n, dt)[...] = (Am.T.astype(np.float64) @ _vec(r, m, dt).astype(np.float64)).astype(T)
        return 0

    def pb_sqdist(self, h, dt, n, x, b, grad):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        d = (_vec(x, n, dt) - _vec(b, n, dt)).astype(T)
        _vec(grad, n, dt)[...] = d
        self._set(L.PB_S_AUX, _fsum_prod(d, d))
        return 0

    # ---- K7: L-BFGS ------------------------------------------------------------------------------------------------
    def pb_lbfgs_create(self, h, dt, n, mem, out):
        key = self._next
        self._next += 1
        self._lb[key] = _Lbfgs(dt, n, mem)
        out._obj.value = key
        return 0

    def pb_lbfgs_destroy(self, op):
        self._lb.pop(_addr(op), None)
        return 0

    def pb_lbfgs_reset(self, op):
        lb = self._lb[_addr(op)]
        lb.currmem = lb.curridx = 0
        lb.H = lb.T(1)
        lb.pending = False
        return 0

    def pb_lbfgs_info(self, op, cm, ci, hh):
        lb = self._lb[_addr(op)]
        cm._obj.value, ci._obj.value, hh._obj.value = lb.currmem, lb.curridx, float(lb.H)
        return 0

    def pb_lbfgs_update(self, h, op, a, ap, b, bp):
        lb = self._lb[_addr(op)]
        T, n, dt = lb.T, lb.n, lb.dt
        self.launches += 1
        s = _vec(a, n, dt).copy() if not _addr(ap) else (_vec(a, n, dt) - _vec(ap, n, dt)).astype(T)
        y = _vec(b, n, dt).copy() if not _addr(bp) else (_vec(b, n, dt) - _vec(bp, n, dt)).astype(T)
        lb.s[lb.spare, :n], lb.y[lb.spare, :n] = s, y
        self._set(L.PB_S_AUX2, _fsum_prod(s, y))
        self._set(L.PB_S_AUX3, _fsum_prod(y, y))
        lb.pending = True
        return 0

    def pb_lbfgs_commit(self, op, ys, yty, accepted):
        lb = self._lb[_addr(op)]
        assert lb.pending
        lb.pending = False
        T = lb.T
        acc = 0
        if T(ys) > 0:
            acc = 1
            lb.ys[lb.spare] = T(ys)
            lb.H = T(T(ys) / T(yty))
            lb.curridx = lb.curridx + 1 if lb.curridx < lb.mem else 1
            lb.currmem = min(lb.currmem + 1, lb.mem)
            ev = lb.slot_of[lb.curridx]
            lb.slot_of[lb.curridx] = lb.spare
            if ev >= 0:
                lb.spare = ev
            else:
                used = {s for s in lb.slot_of[1:] if s >= 0}
                lb.spare = next(k for k in range(lb.mem + 1) if k not in used)
        accepted._obj.value = acc
        return 0

    def pb_lbfgs_apply(self, h, op, v, scale, d, x, x_d):
        lb = self._lb[_addr(op)]
        assert not lb.pending
        T, n, dt = lb.T, lb.n, lb.dt
        dv = _vec(v, n, dt).copy()
        order = []
        idx = lb.curridx
        for _ in range(lb.currmem):
            order.append(lb.slot_of[idx])
            idx = idx - 1 if idx > 1 else lb.mem
        self.launches += 2 * len(order) + 1
        for sl in order:
            lb.alpha[sl] = T(T(_fsum_prod(lb.s[sl, :n], dv)) / lb.ys[sl])
            dv = (dv - (lb.alpha[sl] * lb.y[sl, :n]).astype(T)).astype(T)
        dv = (dv * lb.H).astype(T)
        for sl in reversed(order):
            beta = T(T(_fsum_prod(lb.y[sl, :n], dv)) / lb.ys[sl])
            dv = (dv + (T(lb.alpha[sl] - beta) * lb.s[sl, :n]).astype(T)).astype(T)
        dv = (dv * T(scale)).astype(T)
        _vec(d, n, dt)[...] = dv
        if _addr(x):
            _vec(x_d, n, dt)[...] = (_vec(x, n, dt) + dv).astype(T)
        return 0

    def pb_conj_prox(self, h, dt, n, v, gamma, hd, out):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        hd_ = hd._obj if hasattr(hd, "_obj") else hd
        vv = _vec(v, n, dt).copy()
        if hd_.kind == L.PB_PROX_ZERO:
            _vec(out, n, dt)[...] = 0
            return 0
        p = self._prox(T, hd_, (vv / T(gamma)).astype(T), T(T(1) / T(gamma)), set_gsum=False)
        _vec(out, n, dt)[...] = (vv - (T(gamma) * p).astype(T)).astype(T)
        return 0

    # ---- K9: least-squares prox ---------------------------------------------------------------------------------------
    def pb_lsq_prox_create(self, h, dt, m, n, A, b, lam, out):
        key = self._next
        self._next += 1
        Am = _vec(A, m * n, dt).reshape(n, m).T.astype(np.float64)
        self._lb[key] = ("lsqprox", dt, m, n, Am, _vec(b, m, dt).astype(np.float64), lam)
        out._obj.value = key
        return 0

    def pb_lsq_prox_destroy(self, op):
        self._lb.pop(_addr(op), None)
        return 0

    def pb_lsq_prox_apply(self, h, op, x, gamma, y):
        _, dt, m, n, Am, b, lam = self._lb[_addr(op)]
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 4
        gamma = float(T(gamma))
        xv = _vec(x, n, dt).astype(np.float64)
        yv = np.linalg.solve(lam * Am.T @ Am + np.eye(n) / gamma, lam * Am.T @ b + xv / gamma).astype(T)
        _vec(y, n, dt)[...] = yv
        r = (Am @ yv.astype(np.float64) - b).astype(T)
        self._set(L.PB_S_AUX, _fsum_prod(r, r))
        return 0

    # ---- K8: Douglas-Rachford ----------------------------------------------------------------------------------------
    def pb_dr_step(self, h, dt, n, x, gamma, f, g, x_out, y, r, z, res):
        T = np.float32 if dt == L.PB_F32 else np.float64
        self.launches += 1
        for d in (f, g):
            if (d._obj if hasattr(d, "_obj") else d).kind not in (L.PB_PROX_ZERO, L.PB_PROX_L1, L.PB_PROX_BOX, L.PB_PROX_SQRL2):
                self.err = b"not element-wise"
                return 4
        xv = _vec(x, n, dt).copy()
        yv = self._prox(T, f, xv, gamma, set_gsum=False)
        rv = ((T(2) * yv).astype(T) - xv).astype(T)
        zv = self._prox(T, g, rv, gamma, set_gsum=False)
        sv = (yv - zv).astype(T)
        _vec(x_out, n, dt)[...] = (xv - sv).astype(T)
        for p, val in ((y, yv), (r, rv), (z, zv), (res, sv)):
            if _addr(p):
                _vec(p, n, dt)[...] = val
        self.scal[L.PB_S_RESINF] = float(np.max(np.abs(sv))) if n else 0.0
        return 0


def _tv_step(self, h, dt, H, W, x, b, gamma, lam, x_out, y, z, row0, Hglob, hp, hn):
    """pb_dr_tv_step through the numpy oracle of the splitting (single shard: no halos on the CPU emulation)."""
    from oracle import tv_oracle as tvo

    T = np.float32 if dt == L.PB_F32 else np.float64
    self.launches += 1
    n = H * W
    xv = _vec(x, 5 * n, dt).copy()
    f = tvo.TVSplit(_vec(b, n, dt).copy(), T(lam), (H, W), row0, Hglob)
    yv, _ = f.prox(xv, T(gamma))
    rv = (T(2) * yv - xv).astype(T)
    zt, _ = tvo.Consensus(5).prox(rv, T(gamma))
    res = (yv - zt).astype(T)
    _vec(x_out, 5 * n, dt)[...] = (xv - res).astype(T)
    if _addr(y):
        _vec(y, 5 * n, dt)[...] = yv
    if _addr(z):
        _vec(z, n, dt)[...] = zt[:n]
    self.scal[L.PB_S_RESINF] = float(np.max(np.abs(res))) if n else 0.0
    return 0


EmuLib.pb_dr_tv_step = _tv_step


def _fd_forward(self, h, dt, H, W, u, out):
    from oracle.stencil_oracle import FiniteDifference2D

    self.launches += 1
    _vec(out, 2 * H * W, dt)[...] = FiniteDifference2D(H, W).mul(_vec(u, H * W, dt).copy())
    return 0


def _fd_adjoint(self, h, dt, H, W, pq, out):
    from oracle.stencil_oracle import FiniteDifference2D

    self.launches += 1
    _vec(out, H * W, dt)[...] = FiniteDifference2D(H, W).mul_t(_vec(pq, 2 * H * W, dt).copy())
    return 0


EmuLib.pb_fd2d_forward = _fd_forward
EmuLib.pb_fd2d_adjoint = _fd_adjoint


class EmuContext:
    """Stands in for host.Context: CPU tensors, EmuLib, memcpy-style read-back."""

    def __init__(self):
        self.lib = EmuLib()
        self.h = C.c_void_p(1)
        self.device = torch.device("cpu")
        self.index = 0

    def default_comm(self):
        from proxb200.host import LocalComm

        return LocalComm()

    def read_scalars(self):
        return self.lib.scal.copy()

    def launches(self):
        return self.lib.launches

    def sync(self):
        pass
