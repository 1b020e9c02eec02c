"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Forward-difference operator of anisotropic TV and its adjoint (csrc/stencil_kernels.cu),
for Chambolle-Pock TV denoising (TV itself is not in the reference: parity unpinned, pinned by the
adjoint identity and by agreement with the Douglas-Rachford TV solution of oracle/tv_oracle.py)."""
import numpy as np


class FiniteDifference2D:
    def __init__(self, H, W):
        self.H, self.W = H, W
        self.m, self.n = 2 * H * W, H * W

    def mul(self, u):
        u = u.reshape(self.H, self.W)
        out = np.zeros((2, self.H, self.W), u.dtype)
        out[0, :, :-1] = u[:, 1:] - u[:, :-1]
        out[1, :-1, :] = u[1:, :] - u[:-1, :]
        return out.reshape(-1)

    def mul_t(self, pq):
        T = pq.dtype.type
        p, q = pq.reshape(2, self.H, self.W).copy()
        p[:, -1] = 0
        q[-1, :] = 0
        pl = np.zeros_like(p)
        pl[:, 1:] = p[:, :-1]
        qu = np.zeros_like(q)
        qu[1:, :] = q[:-1, :]
        return ((pl - p).astype(pq.dtype) + (qu - q).astype(pq.dtype)).astype(pq.dtype).reshape(-1)

    def opnorm_bound(self):
        return float(np.sqrt(8.0))
