"""Total-variation denoising by Douglas-Rachford splitting (BASELINE.json configs[4]: 8192 x 8192 image, Float32, 2 x B200).

TV is not part of the reference: the iteration is the reference's DouglasRachford
(src/algorithms/douglas_rachford.jl:54-63), run on the product-space splitting described in csrc/tv_kernels.cu:

    f = TVSplit(b, lam, shape)   separable sum over five stacked copies of the image (data term + four pair sets)
    g = IndConsensus(5)          indicator of {all copies equal}; prox = average

    x0 = TVSplit.initial_point()          # five copies of b
    y, it = DouglasRachford(tol=...)(x0=x0, f=f, g=g, gamma=1.0)
    u = f.image(y)                        # the denoised image

`DouglasRachfordIteration` recognises the pair and runs the whole iteration as ONE kernel pass (K10, `pb_dr_tv_step`).
Row-sharded over several GPUs (one process per GPU): pass `comm=` to TVSplit; each rank holds `shape = (H_local, W)` rows
starting at global row `row0`.  The one image row a shard needs from each neighbour is read by the kernel straight from the
neighbour's memory over NVLink (cudaIpc mapping set up once), and the per-iteration scalar exchange is the only barrier.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .host import Context, LocalComm, real_type, torch


class _RawCuda:
    def __init__(self, addr, n, R):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4" if R is np.float32 else "<f8",
                                         "data": (int(addr), False), "version": 2}


class _IpcBuffer:
    """pb_malloc'ed device memory wrapped as a torch tensor (cudaIpc can only export whole cudaMalloc allocations, which a
    caching-allocator tensor is not)."""

    def __init__(self, ctx, n, R):
        self.ctx = ctx
        p = C.c_void_p()
        nbytes = max(16, int(n) * (4 if R is np.float32 else 8))
        L.check(ctx.lib.pb_malloc(ctx.h, nbytes, C.byref(p)))
        self.ptr = p
        self.tensor = torch().as_tensor(_RawCuda(p.value, n, R), device=ctx.device)

    def handle(self):
        h = C.create_string_buffer(L.PB_IPC_HANDLE_BYTES)
        L.check(self.ctx.lib.pb_ipc_export(self.ctx.h, self.ptr, h))
        return h.raw

    def free(self):
        if self.ptr is not None:
            self.tensor = None
            self.ctx.lib.pb_free(self.ctx.h, self.ptr)
            self.ptr = None


class IndConsensus:
    """Indicator of {x_0 = ... = x_{K-1}} over K stacked copies; prox = the average, broadcast."""

    def __init__(self, K=5):
        self.K = int(K)


class TVSplit:
    """F(X) = 0.5||x_0 - b||^2 + lam * (TV pair sets on x_1..x_4) over five stacked copies of an H x W image (row-major)."""

    tv_split = True
    ncopies = 5

    def __init__(self, b, lam, shape=None, comm=None, row0=0, Hglob=None, device=None):
        t = torch()
        if lam < 0:
            raise ValueError("parameter lambda must be nonnegative")
        if shape is None:
            shape = tuple(b.shape)
        if len(shape) != 2:
            raise ValueError("shape must be (H, W)")
        self.H, self.W = int(shape[0]), int(shape[1])
        ctx = Context.get(device if device is not None else (b.device if isinstance(b, t.Tensor) and b.is_cuda else None))
        self.ctx = ctx
        self.b = (b if isinstance(b, t.Tensor) else t.as_tensor(np.ascontiguousarray(b))).to(ctx.device).contiguous().view(-1)
        if self.b.numel() != self.H * self.W:
            raise ValueError("b does not match shape")
        self.R = real_type(self.b.dtype)
        self.lam = lam
        self.comm = comm or LocalComm()
        self.row0 = int(row0)
        self.Hglob = self.H if Hglob is None else int(Hglob)
        if self.comm.size == 1 and (self.row0 != 0 or self.Hglob != self.H):
            raise ValueError("row0 / Hglob describe a row shard: pass the communicator of the sharded run")

    def initial_point(self):
        """Five copies of b (any starting point works; this one starts at consensus)."""
        return self.b.repeat(self.ncopies)

    def image(self, X):
        """The image carried by a stacked vector (copy 0; at convergence all copies agree)."""
        n = self.H * self.W
        return X[:n].reshape(self.H, self.W)

    def objective(self, u):
        """0.5||u - b||^2 + lam*TV(u) of THIS shard's rows, accumulated in float64 on the host (diagnostics only; the
        vertical differences across a shard boundary are not included)."""
        u64 = np.asarray(u.detach().cpu().numpy() if hasattr(u, "detach") else u, np.float64).reshape(self.H, self.W)
        b64 = self.b.detach().cpu().numpy().astype(np.float64).reshape(self.H, self.W)
        return 0.5 * np.sum((u64 - b64) ** 2) + float(self.lam) * (np.abs(np.diff(u64, axis=1)).sum() + np.abs(np.diff(u64, axis=0)).sum())


def halo_plan(allinfo, rank, Hglob, es):
    """Which neighbour rows a shard reads.  `allinfo`: one (rank, row0, H, W, handle_buf0, handle_buf1) per rank.  Returns
    (prev, next), each None (true image border) or ((handle_buf0, handle_buf1), byte offset of the halo row inside the
    neighbour's [5][H_nb][W] buffer).  The row above a shard is the LAST row of the upper neighbour, in the copy whose vertical pair
    (row0 - 1, row0) it is: copy 3 pairs (even, even + 1), copy 4 pairs (odd, odd + 1); symmetrically below."""
    info = sorted(allinfo, key=lambda r: r[1])
    W = info[0][3]
    start = 0
    for r in info:                                          # shards must tile [0, Hglob) in order
        if r[1] != start or r[3] != W or r[2] <= 0:
            raise ValueError("row shards must be non-empty, contiguous, ordered and of equal width")
        start += r[2]
    if start != Hglob:
        raise ValueError("row shards do not cover Hglob rows")
    pos = [r[0] for r in info].index(rank)
    _, row0, H, _, _, _ = info[pos]
    prev = nxt = None
    if pos > 0:
        _, _, Hp, _, h0, h1 = info[pos - 1]
        kc = 3 if (row0 - 1) % 2 == 0 else 4
        prev = ((h0, h1), (kc * Hp * W + (Hp - 1) * W) * es)
    if pos + 1 < len(info):
        _, _, Hn, _, h0, h1 = info[pos + 1]
        kc = 3 if (row0 + H - 1) % 2 == 0 else 4
        nxt = ((h0, h1), kc * Hn * W * es)
    return prev, nxt


class TVDouglasRachfordEngine:
    """Buffers and neighbour mappings of the fused TV iteration (used by DouglasRachfordIteration)."""

    def __init__(self, f: TVSplit, x0_dev):
        self.f = f
        ctx, R = f.ctx, f.R
        self.ctx = ctx
        n5 = 5 * f.H * f.W
        if x0_dev.numel() != n5:
            raise ValueError(f"x0 must hold 5 stacked copies of the {f.H} x {f.W} image ({n5} entries)")
        comm = f.comm
        self.sharded = comm.size > 1
        if self.sharded:
            self.bufs = [_IpcBuffer(ctx, n5, R), _IpcBuffer(ctx, n5, R)]
            self.X = [b_.tensor for b_ in self.bufs]
        else:
            self.bufs = None
            self.X = [torch().empty_like(x0_dev), torch().empty_like(x0_dev)]
        self.X[0].copy_(x0_dev)
        self.cur = 0
        self.halo = [(None, None), (None, None)]      # per ping-pong buffer: (prev-row pointer, next-row pointer)
        self._opened = []
        if self.sharded:
            self._connect()
            # Safety net for callers that drive the iterator themselves and drop it: when the engine is garbage collected its IPC
            # mappings and pb_malloc'ed buffers (outside torch's caching allocator: ~1.3 GB per rank at 8192^2) are released.  The
            # solver driver does not wait for the collector: it calls `state.release()` (douglas_rachford.py) as soon as the
            # solution has been copied out.
            import weakref

            self._finalizer = weakref.finalize(self, _release_tv_resources, self.ctx, self._opened, self.bufs)

    def _connect(self):
        f, comm = self.f, self.f.comm
        dist = comm.dist
        es = 4 if f.R is np.float32 else 8
        mine = (comm.rank, f.row0, f.H, f.W, self.bufs[0].handle(), self.bufs[1].handle())
        allinfo = [None] * comm.size
        dist.all_gather_object(allinfo, mine, group=getattr(comm, "group", None))
        prev, nxt = halo_plan(allinfo, comm.rank, f.Hglob, es)
        prevs, nexts = [None, None], [None, None]
        for plan, slots in ((prev, prevs), (nxt, nexts)):
            if plan is None:
                continue
            handles, offset = plan
            for i, h in enumerate(handles):
                p = C.c_void_p()
                L.check(self.ctx.lib.pb_ipc_open(self.ctx.h, h, C.byref(p)))
                self._opened.append(p)
                slots[i] = p.value + offset
        self.halo = [(prevs[0], nexts[0]), (prevs[1], nexts[1])]
        torch().cuda.synchronize(self.ctx.device)
        dist.barrier(group=getattr(comm, "group", None))    # every rank's X[0] is filled before anyone reads a halo row

    def step(self, gamma, y=None, z=None, redo=False):
        """One fused iteration X[cur] -> X[1 - cur] (redo: repeat the previous one, to materialise y / z)."""
        f, ctx = self.f, self.ctx
        src = (1 - self.cur) if redo else self.cur
        dst = 1 - src
        hp, hn = self.halo[src]
        L.check(ctx.lib.pb_dr_tv_step(ctx.h, L.PB_F32 if f.R is np.float32 else L.PB_F64, f.H, f.W, C.c_void_p(self.X[src].data_ptr()),
                                      C.c_void_p(f.b.data_ptr()), float(gamma), float(f.R(f.lam)), C.c_void_p(self.X[dst].data_ptr()),
                                      C.c_void_p(y.data_ptr()) if y is not None else None,
                                      C.c_void_p(z.data_ptr()) if z is not None else None, f.row0, f.Hglob,
                                      C.c_void_p(hp) if hp else None, C.c_void_p(hn) if hn else None))
        if not redo:
            self.cur = dst
        return self.X[dst]

    def close(self):
        """Unmap the neighbours' buffers and free this rank's (sharded runs only; idempotent).  `state.x` is a tensor view of these
        buffers: callers that still need it copy it first (`DouglasRachfordState.release` does)."""
        fin = getattr(self, "_finalizer", None)
        if fin is not None:
            fin.detach()
            self._finalizer = None
        _release_tv_resources(self.ctx, getattr(self, "_opened", []), getattr(self, "bufs", None))
        self._opened = []
        self.bufs = None


def _release_tv_resources(ctx, opened, bufs):
    for p in list(opened):
        ctx.lib.pb_ipc_close(ctx.h, p)
    del opened[:]
    if bufs:
        torch().cuda.synchronize(ctx.device)
        for b_ in bufs:
            b_.free()
        del bufs[:]


__all__ = ["TVSplit", "IndConsensus"]
