"""IterationTools: adaptors over the solver iterators (host-side only; no device work of their own).

Reference: src/utilities/iteration_tools.jl -- `halt` (:9-43), `tee` (:47-69), `sample` (:73-109), `stopwatch` (:113-147),
`loop` (:151-160).  Semantics kept: `halt` yields the element that satisfied the predicate and THEN stops; `sample(iter, p)`
yields every p-th element and, for a finite source, the last element of an incomplete final period; `stopwatch` yields
(elapsed nanoseconds since the first `next`, element); `loop` exhausts the iterable and returns the last element.

Device note: the iterators of this package synchronise with the GPU once per element (the scalar read-back of the stop
test), so `stopwatch` measures completed device work, not enqueue time.
"""
from __future__ import annotations

import time


class halt:
    def __init__(self, it, fun):
        self.it, self.fun = it, fun

    def __iter__(self):
        for x in self.it:
            yield x
            if self.fun(x):
                return

    def __len__(self):
        return len(self.it)


class tee:
    def __init__(self, it, fun):
        self.it, self.fun = it, fun

    def __iter__(self):
        for x in self.it:
            self.fun(x)
            yield x

    def __len__(self):
        return len(self.it)


class sample:
    def __init__(self, it, period):
        if int(period) < 1:
            raise ValueError("period must be positive")
        self.it, self.period = it, int(period)

    def __iter__(self):
        k, last, have = 0, None, False
        for x in self.it:
            k += 1
            last, have = x, True
            if k % self.period == 0:
                have = False
                yield x
        if have:
            yield last

    def __len__(self):
        q, r = divmod(len(self.it), self.period)
        return q if r == 0 else q + 1


class stopwatch:
    def __init__(self, it):
        self.it = it

    def __iter__(self):
        t0 = time.perf_counter_ns()
        for x in self.it:
            yield time.perf_counter_ns() - t0, x

    def __len__(self):
        return len(self.it)


def loop(it):
    """Exhaust `it`, return its last element (iteration_tools.jl:151-160; an empty iterable is an error there too)."""
    have = False
    out = None
    for out in it:
        have = True
    if not have:
        raise ValueError("loop: empty iterable")
    return out


__all__ = ["halt", "tee", "sample", "stopwatch", "loop"]
