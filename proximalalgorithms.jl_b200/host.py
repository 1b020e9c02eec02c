"""Host-side plumbing: device context, scalar-block read-back, double-double combination across shards.

torch is used for what the task statement allows it for: device memory (tensors own the iterates), the current stream
and `torch.distributed`.  All arithmetic on the path is done by libproxb200 kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_TORCH = None


def torch():
    global _TORCH
    if _TORCH is None:
        import torch as _t

        _TORCH = _t
    return _TORCH


def real_type(dtype):
    """R = real(eltype(x0)) as a numpy scalar type, from a torch or numpy dtype."""
    s = str(dtype)
    if s.endswith("float32"):
        return np.float32
    if s.endswith("float64"):
        return np.float64
    raise TypeError(f"unsupported element type {dtype}: the B200 path handles Float32 and Float64 vectors")


def pb_dtype(R):
    return L.PB_F32 if R is np.float32 else L.PB_F64


def torch_dtype(R):
    t = torch()
    return t.float32 if R is np.float32 else t.float64


# ---------------------------------------------------------------------------------------------------------------------
# double-double helpers (host mirror of common.cuh): combine per-shard (hi, lo) pairs in rank order, round once
# ---------------------------------------------------------------------------------------------------------------------


def two_sum(a: float, b: float):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def dd_add(a, b):
    s, e = two_sum(a[0], b[0])
    e += a[1] + b[1]
    hi = s + e
    return hi, e - (hi - s)


def _nanmax(vals):
    m = 0.0
    for v in vals:
        if v != v:
            return float("nan")
        if v > m:
            m = v
    return m


class Scalars:
    """Host view of one (combined) scalar block.  Sums are rounded once from their double-double pairs."""

    __slots__ = ("gsum", "res_sq", "gdr", "res_inf", "aux", "aux_inf", "aux2", "aux3", "parts")

    def __init__(self, parts: np.ndarray):
        # parts: (P, PB_NSCALARS) -- one row per shard, in rank order
        self.parts = parts

        def s(slot):
            acc = (0.0, 0.0)
            for p in range(parts.shape[0]):
                acc = dd_add(acc, (float(parts[p, slot]), float(parts[p, slot + 1])))
            return acc[0] + acc[1]

        self.gsum = s(L.PB_S_GSUM)
        self.res_sq = s(L.PB_S_RESSQ)
        self.gdr = s(L.PB_S_GDR)
        self.aux = s(L.PB_S_AUX)
        self.aux2 = s(L.PB_S_AUX2)
        self.aux3 = s(L.PB_S_AUX3)
        self.res_inf = _nanmax(parts[:, L.PB_S_RESINF].tolist())
        self.aux_inf = _nanmax(parts[:, L.PB_S_AUXINF].tolist())


# ---------------------------------------------------------------------------------------------------------------------
# communicator: single device, or one process per GPU with a scalar all-gather per iteration
# ---------------------------------------------------------------------------------------------------------------------


class LocalComm:
    """World of one: the per-iteration exchange degenerates to the pinned read-back of the scalar block."""

    rank = 0
    size = 1

    def exchange(self, ctx) -> Scalars:
        return Scalars(ctx.read_scalars()[None, :])

    def allgather_vector(self, t):
        return t[None, :]


class TorchDistComm:
    """One process per GPU.  The only data-path collective of an elementwise-prox iteration is an all-gather of the
    PB_NSCALARS-double scalar block; the P blocks are then folded in rank order in double-double on every rank, so the
    rounded scalars are identical on all ranks and independent of P."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self._gather = None

    def exchange(self, ctx) -> Scalars:
        t = torch()
        if self._gather is None or self._gather.device != ctx.scal.device:
            self._gather = t.empty((self.size, L.PB_NSCALARS), dtype=t.float64, device=ctx.scal.device)
        self.dist.all_gather_into_tensor(self._gather.view(-1), ctx.scal, group=self.group)
        return Scalars(self._gather.cpu().numpy())

    def allgather_vector(self, v):
        t = torch()
        out = t.empty((self.size, v.numel()), dtype=v.dtype, device=v.device)
        self.dist.all_gather_into_tensor(out.view(-1), v.contiguous(), group=self.group)
        return out


class DeviceExchangeComm:
    """C1 done by the GPU (csrc/xchg.cuh): the last CTA of the fused step pushes this rank's scalar block to every peer
    over NVLink (cudaIpc-mapped buffers), waits for the peers' blocks and drops all rows into mapped pinned host memory;
    `exchange` only polls a pinned flag.  No collective launch, no cudaMemcpy, no stream synchronisation per iteration.
    torch.distributed is used once, to all-gather the 64-byte IPC handles.  World size 1 = zero-copy read-back."""

    def __init__(self, ctx, group=None, fused=True, standalone=False):
        t = torch()
        self.group = group
        self.dist = None
        self.rank, self.size = 0, 1
        self.standalone = standalone
        if not standalone and t.distributed.is_available() and t.distributed.is_initialized():
            self.dist = t.distributed
            self.rank = self.dist.get_rank(group)
            self.size = self.dist.get_world_size(group)
        if self.size > L.PB_MAX_WORLD:
            raise ValueError(f"DeviceExchangeComm supports at most {L.PB_MAX_WORLD} ranks")
        prev = getattr(ctx, "_xchg_comm", None)
        if prev is not None:
            if not prev.standalone:
                raise L.ProxB200Error("this context already owns a device exchange")
            prev.close()          # the implicit single-GPU exchange makes way for an explicit one
        handle = C.create_string_buffer(L.PB_IPC_HANDLE_BYTES)
        L.check(ctx.lib.pb_xchg_init(ctx.h, self.rank, self.size, handle))
        if self.size > 1:
            handles = [None] * self.size
            self.dist.all_gather_object(handles, handle.raw, group=group)
            blob = b"".join(handles)
            L.check(ctx.lib.pb_xchg_connect(ctx.h, C.c_char_p(blob)))
            self.dist.barrier(group=group)
        L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_FUSED_EXCHANGE, 1 if fused else 0))
        self._rows = (C.c_double * (self.size * L.PB_NSCALARS))()
        self.ctx = ctx
        ctx._xchg_comm = self
        self.timeout_s = 20.0

    def exchange(self, ctx) -> Scalars:
        L.check(ctx.lib.pb_exchange_wait(ctx.h, self._rows, self.timeout_s))
        return Scalars(np.frombuffer(self._rows, dtype=np.float64).reshape(self.size, L.PB_NSCALARS).copy())

    def allgather_vector(self, v):
        if self.size == 1:
            return v[None, :]
        t = torch()
        out = t.empty((self.size, v.numel()), dtype=v.dtype, device=v.device)
        self.dist.all_gather_into_tensor(out.view(-1), v.contiguous(), group=self.group)
        return out

    def close(self):
        if self.ctx is not None:
            L.check(self.ctx.lib.pb_ctx_set_option(self.ctx.h, L.PB_OPT_FUSED_EXCHANGE, 0))
            L.check(self.ctx.lib.pb_xchg_shutdown(self.ctx.h))
            self.ctx._xchg_comm = None
            self.ctx = None


def dense_shard_bounds(R, m: int, n: int, size: int):
    """Column ranges [(lo, hi)] of a dense m x n matrix for `size` ranks of the device exchange: boundaries are multiples of the column
    chunk of the residual order (csrc/lsq_order.h), which is what makes the sharded r = A x - b bit-identical to the single-GPU one
    (include/proxb200.h: pb_lsq_dense_residual_sharded).  Ranks may come out empty when there are fewer chunks than ranks."""
    cc = int(L.lib().pb_lsq_dense_chunk_cols(pb_dtype(R), m, n))
    nch = (n + cc - 1) // cc
    base, rem = divmod(nch, size)
    out, c = [], 0
    for r in range(size):
        k = base + (1 if r < rem else 0)
        out.append((min(n, c * cc), min(n, (c + k) * cc)))
        c += k
    return out


def shard_bounds(n: int, size: int, align: int = 32):
    """Contiguous, `align`-element aligned index ranges of an n-vector over `size` ranks (last shard takes the remainder).
    align=32 keeps every fp32 shard 128-byte aligned; NormL21 callers pass a multiple of the group length."""
    if size <= 0:
        raise ValueError("size must be positive")
    units = -(-n // align)
    base, extra = divmod(units, size)
    bounds = [0]
    for r in range(size):
        bounds.append(min(n, bounds[-1] + (base + (1 if r < extra else 0)) * align))
    bounds[-1] = n
    return [(bounds[r], bounds[r + 1]) for r in range(size)]


# ---------------------------------------------------------------------------------------------------------------------
# device context
# ---------------------------------------------------------------------------------------------------------------------


class Context:
    """Wraps a pb_ctx bound to torch's current stream of one CUDA device.  Fails loudly when there is no GPU: no part of
    the path has a CPU implementation in the product."""

    _cache = {}

    @classmethod
    def get(cls, device=None) -> "Context":
        t = torch()
        if not t.cuda.is_available():
            raise L.ProxB200Error("no CUDA device visible: proxb200 has no CPU fallback (the oracle under oracle/ is test-only)")
        idx = t.cuda.current_device() if device is None else t.device(device).index
        if idx is None:
            idx = t.cuda.current_device()
        stream = t.cuda.current_stream(idx).cuda_stream
        key = (idx, stream)
        ctx = cls._cache.get(key)
        if ctx is None:
            ctx = cls(idx, stream)
            cls._cache[key] = ctx
        return ctx

    def __init__(self, index: int, stream: int):
        t = torch()
        self.lib = L.lib()
        self.index = index
        self.device = t.device("cuda", index)
        h = C.c_void_p()
        L.check(self.lib.pb_ctx_create(index, C.c_void_p(stream), 1, C.byref(h)))
        self.h = h
        self.scal = t.zeros(L.PB_NSCALARS, dtype=t.float64, device=self.device)
        L.check(self.lib.pb_ctx_set_scalars_dev(self.h, C.c_void_p(self.scal.data_ptr())))
        self._host = (C.c_double * L.PB_NSCALARS)()

    def default_comm(self):
        """Exchange used when an iteration is given no `comm`: the world-of-one device exchange (zero-copy read-back of the
        scalar block through mapped pinned memory, pushed by the step kernel itself) -- created once per context."""
        cur = getattr(self, "_xchg_comm", None)
        if cur is not None and cur.size == 1:
            return cur
        if cur is not None:
            return LocalComm()    # a multi-rank exchange is attached: an un-sharded solve reads back by memcpy
        return DeviceExchangeComm(self, standalone=True)

    def read_scalars(self) -> np.ndarray:
        L.check(self.lib.pb_read_scalars(self.h, self._host))
        return np.frombuffer(self._host, dtype=np.float64).copy()

    def launches(self) -> int:
        return int(self.lib.pb_ctx_launch_count(self.h))

    def set_launch(self, ctas_per_sm=0, stream_hints=-1, unroll=0, step_impl=0):
        for opt, val in ((L.PB_OPT_CTAS_PER_SM, ctas_per_sm), (L.PB_OPT_STREAM_HINTS, stream_hints), (L.PB_OPT_UNROLL, unroll),
                         (L.PB_OPT_STEP_IMPL, step_impl)):
            L.check(self.lib.pb_ctx_set_option(self.h, opt, val))

    def sync(self):
        L.check(self.lib.pb_ctx_sync(self.h))


def ptr(t_):
    return C.c_void_p(t_.data_ptr()) if t_ is not None else C.c_void_p(0)


def check_vec(t_, n=None, dtype=None):
    if not t_.is_cuda or not t_.is_contiguous() or t_.dim() != 1:
        raise ValueError("device vectors must be contiguous 1-D CUDA tensors")
    if n is not None and t_.numel() != n:
        raise ValueError(f"length mismatch: {t_.numel()} != {n}")
    if dtype is not None and t_.dtype != dtype:
        raise TypeError(f"dtype mismatch: {t_.dtype} != {dtype}")
