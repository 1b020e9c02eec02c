"""Extrapolation-coefficient sequences (host-side scalar recurrences, all arithmetic in R = real(eltype(x0))).

Follows src/accel/nesterov.jl of the reference: FixedNesterovSequence (:1-20), SimpleNesterovSequence (:22-39),
ConstantNesterovSequence (:51-54), AdaptiveNesterovSequence + next! (:56-103).
"""
from __future__ import annotations

import itertools

import numpy as np


class FixedNesterovSequence:
    """beta_k = (t_k - 1)/t_{k+1},  t_{k+1} = (1 + sqrt(1 + 4 t_k^2))/2,  t_1 = 1   (nesterov.jl:14-17)."""

    def __init__(self, R=np.float64):
        self.R = R

    def __iter__(self):
        R = self.R
        t = R(1)
        one, two, four = R(1), R(2), R(4)
        while True:
            t_next = R((one + np.sqrt(R(one + R(four * R(t * t))))) / two)
            yield R(R(t - one) / t_next)
            t = t_next


class SimpleNesterovSequence:
    """beta_k = (k - 1)/(k + 2), k >= 1   (nesterov.jl:36)."""

    def __init__(self, R=np.float64):
        self.R = R

    def __iter__(self):
        R = self.R
        for k in itertools.count(1):
            yield R(R(k - 1) / R(k + 2))


def ConstantNesterovSequence(m, stepsize):
    """repeated((1 - sqrt(m*stepsize))/(1 + sqrt(m*stepsize)))   (nesterov.jl:51-54); m and stepsize share their type."""
    R = type(m) if isinstance(m, (np.float32, np.float64)) else np.float64
    k_inv = R(R(m) * R(stepsize))
    s = np.sqrt(k_inv)
    return itertools.repeat(R(R(R(1) - s) / R(R(1) + s)))


class AdaptiveNesterovSequence:
    """Variable-stepsize theta recurrence (nesterov.jl:56-60, :80); `next(stepsize)` is the reference's `next!` (:89-103)."""

    def __init__(self, m):
        self.R = type(m) if isinstance(m, (np.float32, np.float64)) else np.float64
        self.m = self.R(m)
        self.stepsize = self.R(-1)
        self.theta = self.R(-1)

    def next(self, stepsize):
        R = self.R
        stepsize = R(stepsize)
        if self.stepsize < 0:
            self.stepsize = stepsize
            self.theta = R(np.sqrt(R(self.m * stepsize))) if self.m > 0 else R(1)
        th2 = R(self.theta * self.theta)
        b = R(R(th2 / self.stepsize) - self.m)
        delta = R(R(b * b) + R(R(R(4) * th2) / R(self.stepsize * stepsize)))
        theta = R(R(stepsize * R(np.sqrt(delta) - b)) / R(2))
        beta = R(R(R(stepsize * self.theta) * R(R(1) - self.theta)) / R(R(self.stepsize * theta) + R(stepsize * th2)))
        self.stepsize = stepsize
        self.theta = theta
        return beta

    def peek(self, stepsize):
        """The coefficient `next(stepsize)` would return, without advancing (used to fuse the extrapolation
        speculatively on the adaptive path)."""
        saved = (self.stepsize, self.theta)
        beta = self.next(stepsize)
        self.stepsize, self.theta = saved
        return beta
