"""PANOC (proximal averaged Newton-type method) on the device.

Same names, keyword arguments, defaults, state fields and control flow as the reference (src/algorithms/panoc.jl:41-53
parameters, :57-83 state, :88-112 init, :114-135 direction hooks, :138-255 step, :257-266 stop / solution / display).  Every
n- or m-length vector lives on the device and every pass over one is a libproxb200 kernel.  What differs from a
transliteration, and why:

  * forward step + prox + residual + the three reductions of the FBE are ONE fused pass (K1 `pb_fb_step`, the same kernel
    ForwardBackward uses) with `res` materialised because the quasi-Newton update needs it;
  * the L-BFGS direction `d = -H*res` AND `x_d = x + d` are one asynchronous chain of fused launches with device-resident
    coefficients (csrc/qn_kernels.cu); `update!` is one fused pass whose `x - x_prev`, `res - res_prev` go straight into
    the ring;
  * the reference's `copyto!` of x, Ax, grad_f_Ax, At_grad_f_Ax, z_curr, x_prev, res_prev (panoc.jl:176-177, :189-193) are
    pointer renames: `state.x` and `state.x_d` are the SAME tensor until a line-search backtrack separates them.
    Consequence for readers of the state: `x_prev` / `res_prev` hold the previous iterate / residual (the differences of
    panoc.jl:125-126 are written directly into the L-BFGS ring);
  * with `A = I` (spelled `A=None`) the m-space twins alias their n-space vectors (`Ax is x`, `grad_f_Ax is At_grad_f_Ax`,
    ...) instead of being copies;
  * one host synchronisation per accepted iteration: the scalar block carries f(Ax_d), the FBE reductions, the stop norm
    and the L-BFGS update sums (<s,y>, <y,y>) together; the update kernel is enqueued speculatively behind the step and
    simply re-issued if the line search moves x.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .accel import LBFGS, QuasiNewtonStyle, acceleration_style
from .algorithms import IterativeAlgorithm, _Engine, _resolve, _to_device_copy, f_model
from .functions import Deferred, MatrixOp, Zero
from .host import pb_dtype, ptr, real_type, torch


class PANOCState:
    """panoc.jl:57-83.  Field names are API.  `y` is materialised on demand."""

    def __init__(self):
        self._y = None

    @property
    def y(self):
        """forward point x - gamma*At_grad_f_Ax, recomputed on demand with the same two roundings as the fused kernel."""
        e = self._engine
        if self._y is None:
            self._y = torch().empty_like(self.x)
        L.check(e.lib.pb_forward(e.ctx.h, pb_dtype(self._R), self.x.numel(), ptr(self.x), ptr(self.At_grad_f_Ax),
                                 float(self.gamma), ptr(self._y)))
        return self._y

    @property
    def res_norm_inf(self):
        return self._R(self._sc.res_inf)


def _is_quadratic(f):
    """ProximalCore.is_generalized_quadratic(Tf) (panoc.jl:217): a trait of the type; False unless declared."""
    return bool(getattr(f, "is_generalized_quadratic", False))


class PANOCIteration:
    """panoc.jl:41-53.  `A=None` is the identity; a numpy / torch matrix is wrapped into a device `MatrixOp`."""

    def __init__(self, x0, f=None, A=None, g=None, alpha=0.95, beta=0.5, Lf=None, gamma=None, adaptive=None,
                 minimum_gamma=1e-7, max_backtracks=20, directions=None, comm=None, n_global=None):
        R = real_type(x0.dtype)
        self.R = R
        self.x0 = x0
        self.f = f if f is not None else Zero()
        self.g = g if g is not None else Zero()
        if A is not None and not hasattr(A, "mul_into"):
            A = MatrixOp(A, device=x0.device if hasattr(x0, "is_cuda") and x0.is_cuda else None)
        self.A = A
        self.alpha, self.beta = R(alpha), R(beta)
        self.Lf = Lf
        self.gamma = (None if Lf is None else R(self.alpha / R(Lf))) if gamma is None else gamma     # :49
        self.adaptive = (self.gamma is None) if adaptive is None else bool(adaptive)                  # :50
        self.minimum_gamma = R(minimum_gamma)
        self.max_backtracks = int(max_backtracks)
        self.directions = directions if directions is not None else LBFGS(5)
        self.style = acceleration_style(self.directions)
        self.comm = comm
        self.n_global = n_global     # length of the whole iterate when x0 is this rank's row shard
        self.backtracks = 0          # stepsize halvings (fb_tools.jl:46-55)
        self.tau_backtracks = 0      # line-search halvings (panoc.jl:203-250)

    # ---- helpers ----------------------------------------------------------------------------------------------------
    def _Lc(self, st):
        return self.R(self.alpha / st.gamma)

    def _fbe_from(self, st, sc):
        """f_model(iter, state) + g_z (panoc.jl:85-86, :202) from the step kernel's reductions."""
        return self.R(f_model(self.R, st.f_Ax, sc.gdr, sc.res_sq, self._Lc(st)) + st.g_z)

    def _mul(self, out, x):
        return x if self.A is None else self.A.mul_into(out, x)

    def _mul_t(self, out, v):
        return v if self.A is None else self.A.mul_t_into(out, v)

    def _lincomb(self, e, a, x, b, y, out):
        L.check(e.lib.pb_lincomb2(e.ctx.h, pb_dtype(self.R), out.numel(), float(a), ptr(x), float(b), ptr(y), ptr(out)))

    def _step_kernel(self, st):
        """y, z, res from x and At_grad_f_Ax (:199-201, :246-248): K1 with res materialised."""
        return st._engine.fb_step(self.R, self.g, st.x, st.At_grad_f_Ax, st.gamma, st.z, y_scratch=st._y_scratch, res_out=st.res)

    def _read(self, st, fx, g_of):
        row, sc = st._engine.read()
        st._sc = sc
        st.f_Ax = _resolve(fx, self.R, row, sc)
        st.g_z = g_of(sc)
        return sc

    # ---- init, panoc.jl:88-112 --------------------------------------------------------------------------------------
    def init(self):
        R, t = self.R, torch()
        st = PANOCState()
        e = _Engine(self, self.x0)
        if e.comm.size != 1 and self.A is not None:
            raise L.ProxB200Error("row-sharded PANOC supports A = I only (a matrix A would need its products combined across ranks)")
        st._engine, st._R = e, R
        dt = pb_dtype(R)
        ident = self.A is None
        st.x = _to_device_copy(self.x0, e.ctx)                                              # :89
        n = st.x.numel()
        m = n if ident else self.A.m
        new_n = lambda: t.empty_like(st.x)                        # noqa: E731
        new_m = (lambda: t.empty_like(st.x)) if ident else (lambda: t.empty(m, dtype=st.x.dtype, device=st.x.device))  # noqa: E731
        st.Ax = st.x if ident else self.A.mul_into(new_m(), st.x)                           # :90
        st.grad_f_Ax = new_m()
        fx = e.eval_f(self.f, st.Ax, st.grad_f_Ax)                                          # :91
        if self.gamma is None:                                                              # :92-95, fb_tools.jl:7-12
            row, sc = (e.read() if isinstance(fx, Deferred) else (None, None))
            fx = _resolve(fx, R, row, sc)
            xeps = new_n()
            L.check(e.lib.pb_add_scalar(e.ctx.h, dt, n, ptr(st.x), 1.0, ptr(xeps)))
            Axeps = self._mul(new_m() if not ident else None, xeps)
            geps = new_m()
            e.eval_f(self.f, Axeps, geps)
            L.check(e.lib.pb_sub(e.ctx.h, dt, m, ptr(geps), ptr(st.grad_f_Ax), ptr(geps)))
            if not ident:
                L.check(e.lib.pb_nrm2sq(e.ctx.h, dt, n, ptr(self.A.mul_t_into(xeps, geps))))
            _, sc2 = e.read()
            n_glob = self.n_global if self.n_global is not None else n * e.comm.size
            lower = R(R(np.sqrt(np.float64(sc2.aux))) / R(np.sqrt(np.float64(n_glob))))
            with np.errstate(divide="ignore"):
                st.gamma = R(self.alpha / lower)
        else:
            st.gamma = R(self.gamma)
        st.At_grad_f_Ax = st.grad_f_Ax if ident else self.A.mul_t_into(new_n(), st.grad_f_Ax)   # :96
        st.z, st.res = new_n(), new_n()
        st._y_scratch = None if getattr(self.g, "fused", False) else new_n()
        fx = e.pre_resolve(R, self.g, fx)
        g_of = self._step_kernel(st)                                                        # :97-98, :109
        self._read(st, fx, g_of)
        st.H = self.directions.initialize(st.x, comm=e.comm) if self.style is QuasiNewtonStyle else None   # :110
        st.tau = R(0)
        # work vectors (:69-82).  Pools: the "copies" of the reference are renames between these buffers.
        st.x_prev, st.x_d, st._x_spare = new_n(), new_n(), None
        st._x_pool = [st.x, st.x_prev, st.x_d]
        st.res_prev, st.z_curr, st.d = new_n(), new_n(), new_n()
        st._g_pool = [st.At_grad_f_Ax, new_n()]               # At_grad_f_Ax / At_grad_f_Ax_d
        st.At_grad_f_Ax_d = st._g_pool[1]
        st.At_grad_f_Az = new_n()
        if ident:
            st.Ad, st.Ax_d, st.grad_f_Ax_d, st.Az, st.grad_f_Az = st.d, st.x_d, st.At_grad_f_Ax_d, st.z_curr, st.At_grad_f_Az
        else:
            st.Ad, st.Az, st.grad_f_Az = new_m(), new_m(), new_m()
            st._Ax_pool = [st.Ax, new_m()]
            st._gm_pool = [st.grad_f_Ax, new_m()]
            st.Ax_d, st.grad_f_Ax_d = st._Ax_pool[1], st._gm_pool[1]
        st.f_Ax_d = R(0)
        return st

    # ---- stepsize backtracking with a general A, fb_tools.jl:24-63 as called at panoc.jl:143-159 ---------------------
    def _backtrack_gamma(self, st):
        R, e = self.R, st._engine
        eps = R(np.finfo(R).eps)
        ident = self.A is None
        sc = st._sc
        f_upp = f_model(R, st.f_Ax, sc.gdr, sc.res_sq, self._Lc(st))                        # :42

        def f_at_z():
            if ident:
                st.Az = st.z                                                                # mul!(Az, I, z) is a copy: alias
            else:
                self.A.mul_into(st.Az, st.z)                                                # :43
            fz = e.eval_f(self.f, st.Az, st.grad_f_Az)                                      # :44, :56-58
            row, scf = (e.read() if isinstance(fz, Deferred) else (None, None))
            return _resolve(fz, R, row, scf)

        f_Az = f_at_z()
        tol = R(R(10) * eps * R(R(1) + abs(f_Az)))                                          # :45
        while f_Az > R(f_upp + tol) and st.gamma >= self.minimum_gamma:                     # :46
            st.gamma = R(st.gamma * R(0.5))                                                 # :47
            g_of = self._step_kernel(st)                                                    # :48-50
            _, sc = e.read()
            st._sc = sc
            st.g_z = g_of(sc)
            f_upp = f_model(R, st.f_Ax, sc.gdr, sc.res_sq, self._Lc(st))                    # :51
            f_Az = f_at_z()                                                                 # :52-53
            tol = R(R(10) * eps * R(R(1) + abs(f_Az)))                                      # :54
            self.backtracks += 1
        if st.gamma < self.minimum_gamma:                                                   # :59-61
            import warnings

            warnings.warn(f"stepsize `gamma` became too small ({st.gamma})")
        return f_Az, f_upp

    # ---- step, panoc.jl:138-255 ---------------------------------------------------------------------------------------
    def step(self, st):
        R, e = self.R, st._engine
        dt, n = pb_dtype(R), st.x.numel()
        ident = self.A is None
        inf = R(np.inf)
        f_Az, a, b, c = inf, inf, inf, inf
        if self.adaptive:                                                                   # :141-164
            gamma_prev = st.gamma
            f_Az, f_Az_upp = self._backtrack_gamma(st)
            if st.gamma != gamma_prev and st.H is not None:
                st.H.reset()
        else:
            f_Az_upp = f_model(R, st.f_Ax, st._sc.gdr, st._sc.res_sq, self._Lc(st))         # :166
        FBE_x = R(f_Az_upp + st.g_z)                                                        # :170
        res_sq_x = st._sc.res_sq                                                            # norm(state.res)^2 of :198

        # direction and x_d = x + d (:173, :183); x_prev <- x (:176) is a rename
        xd_buf = next(b_ for b_ in st._x_pool if b_ is not st.x)
        spare = next(b_ for b_ in st._x_pool if b_ is not st.x and b_ is not xd_buf)
        if st.H is not None:
            st.H.mul_into(st.d, st.res, scale=-1.0, x=st.x, x_d=xd_buf)                     # :114-117 + :183
        else:
            L.check(e.lib.pb_scale(e.ctx.h, dt, n, -1.0, ptr(st.res), ptr(st.d)))           # :119-120
            self._lincomb(e, 1.0, st.x, 1.0, st.d, xd_buf)                                  # :183
        st.x_prev, st.x_d, st._x_spare = st.x, xd_buf, spare                                # :176
        st.tau = R(1)                                                                       # :180
        if ident:
            st.Ad, st.Ax_d = st.d, st.x_d
            st.At_grad_f_Ax_d = st._g_pool[0]
            st.grad_f_Ax_d = st.At_grad_f_Ax_d
            fxd = e.eval_f(self.f, st.x_d, st.At_grad_f_Ax_d)                               # :185-187
        else:
            self.A.mul_into(st.Ad, st.d)                                                    # :181
            st.Ax_d = st._Ax_pool[0] if st.Ax is st._Ax_pool[0] else st._Ax_pool[1]
            self._lincomb(e, 1.0, st.Ax, 1.0, st.Ad, st.Ax_d)                               # :184 (in place over Ax)
            st.grad_f_Ax_d = st._gm_pool[0]
            fxd = e.eval_f(self.f, st.Ax_d, st.grad_f_Ax_d)                                 # :185-186
            st.At_grad_f_Ax_d = st._g_pool[0]
            self.A.mul_t_into(st.At_grad_f_Ax_d, st.grad_f_Ax_d)                            # :187
        # :189-194 -- renames instead of copies
        st.x, st.Ax, st.grad_f_Ax, st.At_grad_f_Ax = st.x_d, st.Ax_d, st.grad_f_Ax_d, st.At_grad_f_Ax_d
        st.z_curr, st.z = st.z, st.z_curr
        if ident:
            st.Az = st.z_curr
        st.res_prev, st.res = st.res, st.res_prev                                           # :177
        # :196-198 -- `0.5 / gamma` is a Float64 literal expression in Julia: sigma and threshold are Float64 for R = Float32
        sigma = np.float64(self.beta) * (np.float64(0.5) / np.float64(st.gamma)) * np.float64(R(R(1) - self.alpha))
        tol = R(R(10) * R(np.finfo(R).eps) * R(R(1) + abs(FBE_x)))
        nr = R(np.sqrt(np.float64(res_sq_x)))
        threshold = np.float64(FBE_x) - sigma * np.float64(R(nr * nr)) + np.float64(tol)

        fxd = e.pre_resolve(R, self.g, fxd)
        g_of = self._step_kernel(st)                                                        # :199-201
        if st.H is not None:                                                                # speculative :252 (tau = 1 accepted)
            st.H.enqueue_update(st.x, st.x_prev, st.res, st.res_prev)
        sc = self._read(st, fxd, g_of)
        st.f_Ax_d = st.f_Ax                                                                 # :187, :194
        FBE_new = self._fbe_from(st, sc)                                                    # :202
        st.line_search_trace = (float(FBE_x), float(threshold), float(FBE_new))            # diagnostics: the first test of :205
        moved = False
        for k in range(1, self.max_backtracks + 1):                                         # :204-250
            if np.float64(FBE_new) <= threshold:
                break
            moved = True
            if np.isinf(f_Az) and not ident:                                                # :209-211
                self.A.mul_into(st.Az, st.z_curr)
            st.tau = R(0) if k >= self.max_backtracks else R(st.tau / R(2))                 # :213
            one_m = R(R(1) - st.tau)
            if st.x is st.x_d:                                                              # un-alias before overwriting x
                st.x = st._x_spare
                if ident:
                    st.Ax = st.x
            self._lincomb(e, st.tau, st.x_d, one_m, st.z_curr, st.x)                        # :214
            if not ident:
                if st.Ax is st.Ax_d:
                    st.Ax = st._Ax_pool[1] if st.Ax_d is st._Ax_pool[0] else st._Ax_pool[0]
                self._lincomb(e, st.tau, st.Ax_d, one_m, st.Az, st.Ax)                      # :215
            if st.At_grad_f_Ax is st.At_grad_f_Ax_d:
                st.At_grad_f_Ax = st._g_pool[1] if st.At_grad_f_Ax_d is st._g_pool[0] else st._g_pool[0]
                if ident:
                    st.grad_f_Ax = st.At_grad_f_Ax
            if not ident and st.grad_f_Ax is st.grad_f_Ax_d:
                st.grad_f_Ax = st._gm_pool[1] if st.grad_f_Ax_d is st._gm_pool[0] else st._gm_pool[0]
            if _is_quadratic(self.f):                                                       # :217-237
                if np.isinf(f_Az):
                    fz = e.eval_f(self.f, st.Az, st.grad_f_Az)
                    row, scf = (e.read() if isinstance(fz, Deferred) else (None, None))
                    f_Az = _resolve(fz, R, row, scf)
                if np.isinf(c):
                    if not ident:
                        self.A.mul_t_into(st.At_grad_f_Az, st.grad_f_Az)
                    c = f_Az
                    m_len = st.Ax_d.numel()
                    L.check(e.lib.pb_dot(e.ctx.h, dt, m_len, ptr(st.Ax_d), ptr(st.grad_f_Az)))
                    d1 = R(e.read()[1].aux)
                    L.check(e.lib.pb_dot(e.ctx.h, dt, m_len, ptr(st.Az), ptr(st.grad_f_Az)))
                    d2 = R(e.read()[1].aux)
                    b = R(d1 - d2)
                    a = R(R(st.f_Ax_d - b) - c)
                fx = R(R(R(a * R(st.tau * st.tau)) + R(b * st.tau)) + c)
                if not ident:
                    self._lincomb(e, st.tau, st.grad_f_Ax_d, one_m, st.grad_f_Az, st.grad_f_Ax)
                self._lincomb(e, st.tau, st.At_grad_f_Ax_d, one_m, st.At_grad_f_Az, st.At_grad_f_Ax)
            else:                                                                           # :238-244
                fx = e.pre_resolve(R, self.g, e.eval_f(self.f, st.Ax, st.grad_f_Ax))
                if not ident:
                    self.A.mul_t_into(st.At_grad_f_Ax, st.grad_f_Ax)
            g_of = self._step_kernel(st)                                                    # :246-248
            sc = self._read(st, fx, g_of)
            FBE_new = self._fbe_from(st, sc)                                                # :249
            self.tau_backtracks += 1
        if st.H is not None:                                                                # :252, :123-128
            if moved:
                st.H.enqueue_update(st.x, st.x_prev, st.res, st.res_prev)
                sc = e.read()[1]
            st.H.commit(sc)
        return st

    def __iter__(self):
        st = self.init()
        while True:
            yield st
            st = self.step(st)


# ---------------------------------------------------------------------------------------------------------------------
# defaults (panoc.jl:257-266) and constructor (:268-315)
# ---------------------------------------------------------------------------------------------------------------------


def default_stopping_criterion(tol, it, state):
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(state.res_norm_inf / state.gamma) <= float(tol)


def default_solution(it, state):
    return state.z


def default_display(k, it, state):
    with np.errstate(divide="ignore", invalid="ignore"):
        print("%5d | %.3e | %.3e | %.3e" % (k, float(state.gamma), float(state.res_norm_inf / state.gamma), float(state.tau)))


def _native_panoc(alg, it, tol):
    """Run the whole solve in `pb_panoc_solve` (csrc/panoc_solve.cu) if every ingredient is built in; returns (z, k) or None."""
    import ctypes as C
    import warnings

    from .accel import LBFGS as _LBFGS
    from .accel import NoAcceleration as _NoAcc
    from .algorithms import _like_input

    f, g, R = it.f, it.g, it.R
    if not hasattr(f, "native_descriptor") or not getattr(g, "fused", False):
        return None
    if g.kind not in (L.PB_PROX_ZERO, L.PB_PROX_L1, L.PB_PROX_BOX, L.PB_PROX_L21):
        return None
    if it.A is not None and not isinstance(it.A, MatrixOp):
        return None
    if not isinstance(it.directions, (_LBFGS, _NoAcc)) or (it.gamma is None and not it.adaptive):
        return None
    fdesc = f.native_descriptor()
    if fdesc is None:
        return None
    e = _Engine(it, it.x0)
    if e.comm.size != 1:
        return None
    t = torch()
    x0 = _to_device_copy(it.x0, e.ctx)
    z = t.empty_like(x0)
    n = x0.numel()
    A = it.A
    opts = L.pb_panoc_opts(alg.maxit, float(tol), float(it.alpha), float(it.beta), 0.0 if it.gamma is None else float(R(it.gamma)),
                           float(it.minimum_gamma), 1 if it.adaptive else 0, it.max_backtracks,
                           it.directions.M if isinstance(it.directions, _LBFGS) else 0, 1 if _is_quadratic(f) else 0,
                           A.m if A is not None else 0, A.n if A is not None else 0, A.A_cm.data_ptr() if A is not None else None)
    gdesc = g.descriptor(R)
    res = L.pb_panoc_result()
    L.check(e.lib.pb_panoc_solve(e.ctx.h, pb_dtype(R), n, C.byref(fdesc), C.byref(gdesc), C.byref(opts), ptr(x0), ptr(z), C.byref(res)))
    if res.warned_small_gamma:
        warnings.warn(f"stepsize `gamma` became too small ({R(res.gamma)})")
    it.backtracks, it.tau_backtracks = int(res.gamma_backtracks), int(res.tau_backtracks)
    alg.last_native = res
    return _like_input(it.x0, z), int(res.iterations)


class _PanocAlgorithm(IterativeAlgorithm):
    """driver="native": the loop runs inside the library (pb_panoc_solve) when all ingredients are built in; "python" (default while
    the native twin awaits its first hardware run) keeps the loop in this file."""

    def __call__(self, **kwargs):
        if self.driver == "native":
            it = self.iterator_type(**{**self.kwargs, **kwargs})
            tol = getattr(self.stop, "_default_tol", None)
            out = None
            if tol is not None and self.solution is default_solution and not self.verbose:
                out = _native_panoc(self, it, tol)
            if out is None:
                raise L.ProxB200Error("driver='native' requested but this PANOC problem needs the Python loop")
            self.last_driver = "native"
            self.last_iteration = it
            return out
        self.last_driver = "python"
        saved, self.driver = self.driver, "python"
        try:
            return super().__call__(**kwargs)
        finally:
            self.driver = saved


def PANOC(maxit=1_000, tol=1e-8, stop=None, solution=default_solution, verbose=False, freq=10, display=default_display,
          driver="python", **kwargs):
    """panoc.jl:296-315."""
    if stop is None:
        def stop(it, state, _tol=tol):
            return default_stopping_criterion(_tol, it, state)
        stop._default_tol = tol
    return _PanocAlgorithm(PANOCIteration, maxit, stop, solution, verbose, freq, display, driver=driver, **kwargs)

