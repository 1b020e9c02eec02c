"""First-order oracles of the hot path: smooth terms (value_and_gradient) and proximable terms (prox!).

Mirrors the callback contract of the reference:
  * `value_and_gradient(f, x) -> (f(x), grad)`   src/ProximalAlgorithms.jl:27-40, benchmark/benchmarks.jl:11-28
  * `prox!(z, g, y, gamma) -> g(z)`              ProximalCore; call sites fast_forward_backward.jl:80,141,
                                                  forward_backward.jl:72,118, fb_tools.jl:49
User types plug in by duck typing: any object with `value_and_gradient(x)` is a smooth term, any object with
`prox_(z, y, gamma)` (in place, returns g(z)) is a proximable term; both operate on 1-D CUDA tensors.  The built-in
types below additionally expose the device fast path (kernels of libproxb200 with the value left in the scalar block).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .host import Context, DeviceExchangeComm, LocalComm, check_vec, pb_dtype, ptr, real_type, torch


class Deferred:
    """A scalar whose value is still in the device scalar block; resolved after the per-iteration read-back.
    `fn(local_row, combined)` gets this rank's raw row and the rank-combined `Scalars`."""

    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn

    def resolve(self, local_row, combined):
        return self.fn(local_row, combined)


def _sq_half(R, sum_sq):
    """`norm(v)^2 / 2` with Julia's sqrt-then-square rounding (benchmark/benchmarks.jl:16) from an exact sum of squares."""
    nr = R(np.sqrt(np.float64(sum_sq)))
    return R(R(nr * nr) / R(2))


def _as_device(a, device, dtype=None):
    t = torch()
    if isinstance(a, t.Tensor):
        out = a.to(device=device, dtype=dtype if dtype is not None else a.dtype)
    else:
        out = t.as_tensor(np.ascontiguousarray(a)).to(device=device, dtype=dtype)
    return out


# =====================================================================================================================
# smooth terms
# =====================================================================================================================


class Zero:
    """ProximalCore.Zero: f = 0 with gradient zero(x) (src/ProximalAlgorithms.jl:38-40); as g its prox is the identity."""

    kind = L.PB_PROX_ZERO
    fused = True

    def value_and_gradient(self, x):
        return real_type(x.dtype)(0), torch().zeros_like(x)

    def value_and_gradient_into(self, ctx, x, grad):
        grad.zero_()
        return real_type(x.dtype)(0)

    def descriptor(self, R):
        return L.pb_prox(L.PB_PROX_ZERO, 0, 0.0, 0.0, None, None)

    def value_from(self, R, scal):
        return R(0)


class MatrixOp:
    """Dense linear map `A` of the composite problems f(Ax) + g(x) (panoc.jl:43; `A = I` is spelled `A=None`).
    Column-major on the device; `A*x` and `A'*v` are the K4 GEMV kernels (the same ones LeastSquares uses)."""

    def __init__(self, A, device=None):
        t = torch()
        ctx = Context.get(device)
        self.ctx = ctx
        if isinstance(A, t.Tensor):
            self.R = real_type(A.dtype)
            self.A_cm = A.to(ctx.device).t().contiguous()
        else:
            A = np.asarray(A)
            self.R = real_type(A.dtype)
            self.A_cm = t.as_tensor(np.ascontiguousarray(A.T)).to(ctx.device)
        self.m, self.n = int(self.A_cm.shape[1]), int(self.A_cm.shape[0])
        self.dtype = self.A_cm.dtype

    def mul_into(self, out, x):
        """out = A * x (also leaves ||A x||^2 in AUX)."""
        check_vec(x, self.n, self.dtype)
        check_vec(out, self.m, self.dtype)
        ctx = self.ctx
        L.check(ctx.lib.pb_lsq_dense_residual(ctx.h, pb_dtype(self.R), self.m, self.n, ptr(self.A_cm), self.m, ptr(x), None, ptr(out)))
        return out

    def mul_t_into(self, out, v):
        """out = A' * v."""
        check_vec(v, self.m, self.dtype)
        check_vec(out, self.n, self.dtype)
        ctx = self.ctx
        L.check(ctx.lib.pb_lsq_dense_gradient(ctx.h, pb_dtype(self.R), self.m, self.n, ptr(self.A_cm), self.m, ptr(v), ptr(out)))
        return out


class FiniteDifference2D:
    """Forward differences of an H x W row-major image, L u = (D_x u, D_y u) in R^{2HW} (zero in the last column / row), and
    the exact adjoint: the linear map of anisotropic total variation lambda*||L u||_1 for the primal-dual methods
    (`ChambollePock(g=SqrNormL2(1, b), h=NormL1(lam), L=FiniteDifference2D(H, W))`)."""

    def __init__(self, H, W, device=None):
        self.ctx = Context.get(device)
        self.H, self.W = int(H), int(W)
        self.n, self.m = self.H * self.W, 2 * self.H * self.W

    def _dt(self, x):
        return pb_dtype(real_type(x.dtype))

    def mul_into(self, out, x):
        check_vec(x, self.n)
        check_vec(out, self.m, x.dtype)
        L.check(self.ctx.lib.pb_fd2d_forward(self.ctx.h, self._dt(x), self.H, self.W, ptr(x), ptr(out)))
        return out

    def mul_t_into(self, out, v):
        check_vec(v, self.m)
        check_vec(out, self.n, v.dtype)
        L.check(self.ctx.lib.pb_fd2d_adjoint(self.ctx.h, self._dt(v), self.H, self.W, ptr(v), ptr(out)))
        return out

    def opnorm(self):
        """Upper bound sqrt(8) of ||L|| (attained in the limit of large images); the default primal-dual stepsizes only need a bound."""
        return float(np.sqrt(8.0))


class LeastSquares:
    """f(x) = 0.5*||A x - b||^2 for a dense matrix, with the benchmark's explicit gradient
    `res = A*x - b; (norm(res)^2/2, A'*res)` (benchmark/benchmarks.jl:11-17).

    A is kept column-major on the device (Julia layout).  With `comm` of size P > 1, `A` is this rank's COLUMN shard
    (matching the row shard of x) and `A_p' r` is local.  Two forms of the exchange:
      * `comm` is a DeviceExchangeComm and `n_global`, `col_offset` are given (shards from `host.dense_shard_bounds`): ONE kernel
        combines the rank's chunk partials, pushes them to the peers over NVLink and folds all chunks in global chunk order
        (csrc/lsq_kernels.cu: k_gemv_n_combine_x) -- r, f and the whole solve are bit-identical to one GPU, and the native
        driver loop (pb_solve) runs sharded;
      * otherwise the m-vector of partial products is all-gathered with NCCL and folded in rank order (deterministic, every rank
        the same, but not the single-GPU summation order).
    """

    is_generalized_quadratic = True      # ProximalOperators' trait of LeastSquares, read by panoc.jl:217

    def __init__(self, A, b, comm=None, device=None, n_global=None, col_offset=None):
        t = torch()
        ctx = Context.get(device)
        self.ctx = ctx
        self.comm = comm or LocalComm()
        self.n_global, self.col_offset = n_global, col_offset
        self.fused_gather = self.comm.size > 1 and n_global is not None and col_offset is not None and isinstance(self.comm, DeviceExchangeComm)
        if isinstance(A, t.Tensor):
            R = real_type(A.dtype)
            a_cm = A.to(ctx.device).t().contiguous()          # (n, m) row-major == A column-major
        else:
            A = np.asarray(A)
            R = real_type(A.dtype)
            a_cm = t.as_tensor(np.ascontiguousarray(A.T)).to(ctx.device)
        self.R = R
        self.m, self.n = int(a_cm.shape[1]), int(a_cm.shape[0])
        self.A_cm = a_cm
        self.b = _as_device(b, ctx.device, a_cm.dtype).contiguous()
        check_vec(self.b, self.m, a_cm.dtype)
        self.r = t.empty(self.m, dtype=a_cm.dtype, device=ctx.device)
        self.calls = 0
        if self.fused_gather:
            # the landing zone of the chunk partials is a fixed region of the exchange buffer (csrc/xchg.cuh: PB_XCHG_VEC_BYTES = 128 MB,
            # [parity][global chunk][row] 8-byte words, two per double); a matrix that does not fit takes the NCCL all-gather instead --
            # the same decision on every rank, since it depends on (m, n_global, dtype) only
            cc = int(ctx.lib.pb_lsq_dense_chunk_cols(pb_dtype(R), self.m, int(n_global)))
            nch = (int(n_global) + cc - 1) // cc
            if 2 * nch * self.m * (np.dtype(R).itemsize // 4) * 8 > (128 << 20):
                self.fused_gather = False

    def _residual(self, x):
        lib, ctx, dt = self.ctx.lib, self.ctx, pb_dtype(self.R)
        check_vec(x, self.n, self.A_cm.dtype)
        if self.comm.size == 1:
            L.check(lib.pb_lsq_dense_residual(ctx.h, dt, self.m, self.n, ptr(self.A_cm), self.m, ptr(x), ptr(self.b), ptr(self.r)))
            return Deferred(lambda row, comb: _sq_half(self.R, row[L.PB_S_AUX] + row[L.PB_S_AUX + 1]))
        if self.fused_gather:
            L.check(lib.pb_lsq_dense_residual_sharded(ctx.h, dt, self.m, self.n, ptr(self.A_cm) if self.n else None, self.m, ptr(x) if self.n else None,
                                                      ptr(self.b), ptr(self.r), int(self.n_global), int(self.col_offset), 0))
            return Deferred(lambda row, comb: _sq_half(self.R, row[L.PB_S_AUX] + row[L.PB_S_AUX + 1]))
        # column shard: partial product, all-gather, fixed-order fold, subtract b, ||r||^2 (replicated on every rank)
        L.check(lib.pb_lsq_dense_residual(ctx.h, dt, self.m, self.n, ptr(self.A_cm), self.m, ptr(x), None, ptr(self.r)))
        parts = self.comm.allgather_vector(self.r)
        acc = parts[0].clone()
        for p in range(1, parts.shape[0]):
            acc += parts[p]
        L.check(lib.pb_sub(ctx.h, dt, self.m, ptr(acc), ptr(self.b), ptr(self.r)))
        return Deferred(lambda row, comb: _sq_half(self.R, row[L.PB_S_AUX] + row[L.PB_S_AUX + 1]))

    def value_and_gradient_into(self, ctx, x, grad):
        """Device fast path: enqueue r = A x - b (AUX = ||r||^2) and grad = A' r; the value is Deferred."""
        self.calls += 1
        val = self._residual(x)
        L.check(ctx.lib.pb_lsq_dense_gradient(ctx.h, pb_dtype(self.R), self.m, self.n, ptr(self.A_cm), self.m, ptr(self.r), ptr(grad)))
        return val

    def value_into(self, ctx, x):
        """f(x) only: the reference discards the gradient of the FFB line search (fast_forward_backward.jl:112-128
        passes grad_f_Az = nothing), so skipping A' r there is parity-safe (hard part 8 iii)."""
        self.calls += 1
        return self._residual(x)

    # ---- as a PROXIMABLE term (DouglasRachford's f, test/problems/test_lasso_small.jl:39,205-214) ---------------------
    def prox_enqueue(self, ctx, y, x, gamma, comm=None):
        """prox!(y, f, x, gamma) of ProximalOperators' LeastSquaresDirect (K9, csrc/lsq_prox.cu); AUX = ||A y - b||^2."""
        if self.comm.size != 1:
            raise L.ProxB200Error("the factorised least-squares prox is single-GPU")
        if getattr(self, "_prox_h", None) is None:
            h = C.c_void_p()
            L.check(ctx.lib.pb_lsq_prox_create(ctx.h, pb_dtype(self.R), self.m, self.n, ptr(self.A_cm), ptr(self.b), 1.0, C.byref(h)))
            self._prox_h = h
        L.check(ctx.lib.pb_lsq_prox_apply(ctx.h, self._prox_h, ptr(x), float(self.R(gamma)), ptr(y)))

    def prox_(self, y, x, gamma):
        self.prox_enqueue(self.ctx, y, x, gamma)
        row = self.ctx.read_scalars()
        return _sq_half(self.R, row[L.PB_S_AUX] + row[L.PB_S_AUX + 1])

    def __del__(self):
        h, self._prox_h = getattr(self, "_prox_h", None), None
        if h is not None:
            try:
                self.ctx.lib.pb_lsq_prox_destroy(h)
            except Exception:
                pass

    def native_descriptor(self):
        """pb_smooth for the native driver; a column shard qualifies when its exchange is the in-kernel one (fused_gather)."""
        if self.comm.size != 1:
            if not self.fused_gather:
                return None
            return L.pb_smooth(L.PB_F_LSQ_DENSE, 0, self.m, self.n, self.m, int(self.col_offset), 0, int(self.n_global),
                               self.A_cm.data_ptr() if self.n else 0, self.b.data_ptr(), self.r.data_ptr())
        return L.pb_smooth(L.PB_F_LSQ_DENSE, 0, self.m, self.n, self.m, 0, 0, 0, self.A_cm.data_ptr(), self.b.data_ptr(), self.r.data_ptr())

    def value_and_gradient(self, x):
        """Reference-shaped allocating form."""
        grad = torch().empty_like(x)
        val = self.value_and_gradient_into(self.ctx, x, grad)
        row = self.ctx.read_scalars()
        return val.resolve(row, None), grad


class BlockDiagLeastSquares:
    """f(x) = 0.5*||A x - b||^2, A = blockdiag(A_1..A_B), A_k in R^{mb x nb} ("A implicit, Ax via batched GEMV",
    BASELINE.json configs[1]; structure fixed by SURVEY.md section 7 hard part 6).  `blocks_cm` is a (B, nb, mb) tensor, i.e.
    every block column-major.  Sharding: whole blocks per rank, no vector collective; the value is summed with the
    scalar block."""

    is_generalized_quadratic = True

    def __init__(self, blocks_cm, b, comm=None):
        t = torch()
        if not (isinstance(blocks_cm, t.Tensor) and blocks_cm.is_cuda and blocks_cm.dim() == 3 and blocks_cm.is_contiguous()):
            raise ValueError("blocks_cm must be a contiguous (B, nb, mb) CUDA tensor (each block column-major)")
        self.ctx = Context.get(blocks_cm.device)
        self.comm = comm or LocalComm()
        self.R = real_type(blocks_cm.dtype)
        self.nblk, self.nb, self.mb = (int(s) for s in blocks_cm.shape)
        self.A = blocks_cm
        self.b = b
        check_vec(b, self.nblk * self.mb, blocks_cm.dtype)
        self.r = t.empty_like(b)
        self.n = self.nblk * self.nb
        self.calls = 0

    @classmethod
    def from_numpy(cls, blocks, b, device=None, comm=None):
        """blocks: (B, mb, nb) array in the mathematical orientation."""
        t = torch()
        ctx = Context.get(device)
        cm = t.as_tensor(np.ascontiguousarray(np.transpose(blocks, (0, 2, 1)))).to(ctx.device)
        return cls(cm, t.as_tensor(np.ascontiguousarray(b)).to(ctx.device), comm=comm)

    def native_descriptor(self):
        return L.pb_smooth(L.PB_F_LSQ_BLOCKDIAG, 0, 0, 0, 0, self.nblk, self.mb, self.nb, self.A.data_ptr(), self.b.data_ptr(), self.r.data_ptr())

    def _residual(self, ctx, x):
        check_vec(x, self.n, self.A.dtype)
        L.check(ctx.lib.pb_lsq_blockdiag_residual(ctx.h, pb_dtype(self.R), self.nblk, self.mb, self.nb, ptr(self.A), ptr(x), ptr(self.b), ptr(self.r)))
        return Deferred(lambda row, comb: _sq_half(self.R, comb.aux if comb is not None else row[L.PB_S_AUX] + row[L.PB_S_AUX + 1]))

    def value_and_gradient_into(self, ctx, x, grad):
        self.calls += 1
        check_vec(x, self.n, self.A.dtype)
        # one call: for matrices beyond L2 the library reads every block from HBM once and sweeps it again from L2 (csrc/lsq_fused.cu)
        L.check(ctx.lib.pb_lsq_blockdiag_value_and_gradient(ctx.h, pb_dtype(self.R), self.nblk, self.mb, self.nb, ptr(self.A), ptr(x), ptr(self.b),
                                                            ptr(self.r), ptr(grad)))
        return Deferred(lambda row, comb: _sq_half(self.R, comb.aux if comb is not None else row[L.PB_S_AUX] + row[L.PB_S_AUX + 1]))

    def value_into(self, ctx, x):
        self.calls += 1
        return self._residual(ctx, x)

    def value_and_gradient(self, x):
        grad = torch().empty_like(x)
        val = self.value_and_gradient_into(self.ctx, x, grad)
        return val.resolve(self.ctx.read_scalars(), None), grad


class SquaredDistance:
    """benchmark/benchmarks.jl:19-28: f(x) = norm(x - b)^2/2 with gradient x - b (one fused pass)."""

    def __init__(self, b, device=None):
        self.ctx = Context.get(device if device is not None else (b.device if hasattr(b, "is_cuda") and b.is_cuda else None))
        self.b = _as_device(b, self.ctx.device).contiguous()
        self.R = real_type(self.b.dtype)

    def native_descriptor(self):
        return L.pb_smooth(L.PB_F_SQDIST, 0, 0, 0, 0, 0, 0, 0, None, self.b.data_ptr(), None)

    def value_and_gradient_into(self, ctx, x, grad):
        check_vec(x, self.b.numel(), self.b.dtype)
        L.check(ctx.lib.pb_sqdist(ctx.h, pb_dtype(self.R), x.numel(), ptr(x), ptr(self.b), ptr(grad)))
        return Deferred(lambda row, comb: _sq_half(self.R, comb.aux if comb is not None else row[L.PB_S_AUX] + row[L.PB_S_AUX + 1]))

    def value_and_gradient(self, x):
        grad = torch().empty_like(x)
        val = self.value_and_gradient_into(self.ctx, x, grad)
        return val.resolve(self.ctx.read_scalars(), None), grad


class LinearFunction:
    """f(x) = <c, x>: constant gradient c.  This is the "gradient supplied as a buffer" smooth term of the fused-step-only
    workloads (BASELINE.json configs[2]): no copy is made when the iteration's gradient buffer IS
    `c` (`gradient_buffer`)."""

    def __init__(self, c):
        self.c = c
        self.R = real_type(c.dtype)
        self.ctx = Context.get(c.device)

    def gradient_buffer(self):
        return self.c

    def native_descriptor(self):
        return L.pb_smooth(L.PB_F_LINEAR, 0, 0, 0, 0, 0, 0, 0, None, self.c.data_ptr(), None)

    def value_and_gradient_into(self, ctx, x, grad):
        if grad.data_ptr() != self.c.data_ptr():
            grad.copy_(self.c)
        L.check(ctx.lib.pb_dot(ctx.h, pb_dtype(self.R), x.numel(), ptr(self.c), ptr(x)))
        return Deferred(lambda row, comb: self.R(comb.aux if comb is not None else row[L.PB_S_AUX] + row[L.PB_S_AUX + 1]))

    def value_and_gradient(self, x):
        val = self.value_and_gradient_into(self.ctx, x, self.c)
        return val.resolve(self.ctx.read_scalars(), None), self.c.clone()


# =====================================================================================================================
# proximable terms (ProximalOperators.jl 0.15 semantics; the kernels live in csrc/step_kernels.cu)
# =====================================================================================================================


class _FusedProx:
    fused = True

    def prox_enqueue(self, ctx, z, y, gamma, comm=None):
        """prox!(z, g, y, gamma) with the value left in the scalar block (GSUM): no host synchronisation."""
        R = real_type(y.dtype)
        d = self.descriptor(R)
        L.check(ctx.lib.pb_prox_apply(ctx.h, pb_dtype(R), y.numel(), ptr(y), float(gamma), C.byref(d), ptr(z)))

    def prox_(self, z, y, gamma):
        """Reference-shaped in-place prox!(z, g, y, gamma) -> g(z) (standalone kernel K3)."""
        ctx = Context.get(y.device)
        R = real_type(y.dtype)
        self.prox_enqueue(ctx, z, y, gamma)
        if self.kind in (L.PB_PROX_L1, L.PB_PROX_L21, L.PB_PROX_SQRL2):
            row = ctx.read_scalars()
            return self.value_from(R, row[L.PB_S_GSUM] + row[L.PB_S_GSUM + 1])
        return R(0)

    def value_from(self, R, gsum):
        return R(0)


class NormL1(_FusedProx):
    """g(x) = lambda*||x||_1 (benchmark/benchmarks.jl:52,60)."""

    kind = L.PB_PROX_L1

    def __init__(self, lam=1.0):
        if lam < 0:
            raise ValueError("parameter lambda must be nonnegative")
        self.lam = lam

    def descriptor(self, R):
        return L.pb_prox(L.PB_PROX_L1, 0, float(R(self.lam)), 0.0, None, None)

    def value_from(self, R, gsum):
        return R(R(self.lam) * R(gsum))


class IndBox(_FusedProx):
    """Indicator of [lo, hi] (scalars or per-element CUDA vectors); prox = clamp (test/problems/test_nonconvex_qp.jl:19,33)."""

    kind = L.PB_PROX_BOX

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def descriptor(self, R):
        t = torch()
        lo_v = self.lo if isinstance(self.lo, t.Tensor) else None
        hi_v = self.hi if isinstance(self.hi, t.Tensor) else None
        return L.pb_prox(
            L.PB_PROX_BOX, 0,
            0.0 if lo_v is not None else float(R(self.lo)),
            0.0 if hi_v is not None else float(R(self.hi)),
            lo_v.data_ptr() if lo_v is not None else None,
            hi_v.data_ptr() if hi_v is not None else None,
        )


class NormL21(_FusedProx):
    """g(X) = lambda * sum_j ||X[:, j]||_2 over contiguous groups of `group` entries (NormL21(lambda, 1) on a
    group x ngroups column-major matrix).  Shards must hold whole groups."""

    kind = L.PB_PROX_L21

    def __init__(self, lam=1.0, group=128):
        if lam < 0:
            raise ValueError("parameter lambda must be nonnegative")
        self.lam, self.group = lam, int(group)

    def descriptor(self, R):
        return L.pb_prox(L.PB_PROX_L21, self.group, float(R(self.lam)), 0.0, None, None)

    def value_from(self, R, gsum):
        return R(R(self.lam) * R(gsum))


class SqrNormL2(_FusedProx):
    """f(x) = lam/2 * ||x - b||^2: ProximalOperators' `Translate(SqrNormL2(lam), -b)` (test/problems/test_lasso_small.jl:38;
    b=None is the plain SqrNormL2).  Both a smooth term (value_and_gradient) and an element-wise proximable term
    (prox z = (y - b)/(1 + gamma*lam) + b), so it can sit on either side of a splitting."""

    kind = L.PB_PROX_SQRL2
    is_generalized_quadratic = True

    def __init__(self, lam=1.0, b=None, device=None):
        if lam < 0:
            raise ValueError("parameter lambda must be nonnegative")
        self.lam = lam
        self.b = None
        if b is not None:
            ctx = Context.get(device if device is not None else (b.device if hasattr(b, "is_cuda") and b.is_cuda else None))
            self.b = _as_device(b, ctx.device).contiguous()

    def descriptor(self, R):
        return L.pb_prox(L.PB_PROX_SQRL2, 0, float(R(self.lam)), 0.0, self.b.data_ptr() if self.b is not None else None, None)

    def value_from(self, R, gsum):
        return R(R(R(self.lam) / R(2)) * R(gsum))

    def value_and_gradient_into(self, ctx, x, grad):
        """grad = lam*(x - b), value = lam/2*||x - b||^2 (sqrt-then-square like `norm(.)^2`)."""
        R = real_type(x.dtype)
        dt, n = pb_dtype(R), x.numel()
        if self.b is not None:
            L.check(ctx.lib.pb_sqdist(ctx.h, dt, n, ptr(x), ptr(self.b), ptr(grad)))
        else:
            L.check(ctx.lib.pb_scale(ctx.h, dt, n, 1.0, ptr(x), ptr(grad)))
            L.check(ctx.lib.pb_nrm2sq(ctx.h, dt, n, ptr(x)))
        if float(R(self.lam)) != 1.0:
            L.check(ctx.lib.pb_scale(ctx.h, dt, n, float(R(self.lam)), ptr(grad), ptr(grad)))
        lam = R(self.lam)

        def val(row, comb):
            nr = R(np.sqrt(np.float64(comb.aux if comb is not None else row[L.PB_S_AUX] + row[L.PB_S_AUX + 1])))
            return R(R(lam / R(2)) * R(nr * nr))

        return Deferred(val)

    def value_and_gradient(self, x):
        ctx = Context.get(x.device)
        grad = torch().empty_like(x)
        val = self.value_and_gradient_into(ctx, x, grad)
        return val.resolve(ctx.read_scalars(), None), grad


class IndBallL2:
    """Indicator of {||x||_2 <= r}.  Two-phase prox (global norm, then scale): not single-pass fusable
    (SURVEY.md section 7 hard part 7).  Phase 1 = pb_forward / pb_nrm2sq (AUX = ||y||^2, combined across shards), phase 2 = the
    fused step with PB_PROX_SCALE."""

    fused = False
    kind = L.PB_PROX_SCALE

    def __init__(self, r=1.0):
        if r <= 0:
            raise ValueError("parameter r must be positive")
        self.r = r

    def scale_factor(self, R, ysq):
        ny = R(np.sqrt(np.float64(ysq)))
        with np.errstate(divide="ignore"):
            return R(R(self.r) / ny)          # scal > 1  <=>  y is inside the ball: kernel copies y

    def scale_descriptor(self, R, ysq):
        return L.pb_prox(L.PB_PROX_SCALE, 0, float(self.scale_factor(R, ysq)), 0.0, None, None)

    def ball_descriptor(self, R):
        """One GPU: PB_PROX_BALL -- the step enqueues the norm pass itself and forms r/||y|| on the device (no host round trip)."""
        return L.pb_prox(L.PB_PROX_BALL, 0, float(R(self.r)), 0.0, None, None)

    def prox_enqueue(self, ctx, z, y, gamma, comm=None):
        R = real_type(y.dtype)
        L.check(ctx.lib.pb_nrm2sq(ctx.h, pb_dtype(R), y.numel(), ptr(y)))
        sc = (comm or LocalComm()).exchange(ctx)
        d = self.scale_descriptor(R, sc.aux)
        L.check(ctx.lib.pb_prox_apply(ctx.h, pb_dtype(R), y.numel(), ptr(y), float(gamma), C.byref(d), ptr(z)))

    def prox_(self, z, y, gamma, comm=None):
        self.prox_enqueue(Context.get(y.device), z, y, gamma, comm=comm)
        return real_type(y.dtype)(0)

    def value_from(self, R, gsum):
        return R(0)
