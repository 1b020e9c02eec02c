"""test/utilities/test_iteration_tools.jl restated for the host-side adaptors (CPU only)."""
import itertools
import time

import numpy as np
import pytest

from proxb200 import iteration_tools as IT


def fibonacci(s0=0, s1=1):
    a, b = s0, s1
    while True:
        yield a
        a, b = b, a + b


def test_looping():
    rng = np.random.default_rng(0)
    it = rng.random(10)
    assert IT.loop(it) == it[-1]
    with pytest.raises(ValueError):
        IT.loop([])


def test_halting():
    fib = list(itertools.islice(fibonacci(), 21))
    truncated = IT.halt(fib, lambda x: x >= 1000)
    assert len(truncated) == len(fib)
    assert IT.loop(truncated) == 1597
    assert list(IT.halt(fibonacci(), lambda x: x >= 5)) == [0, 1, 1, 2, 3, 5]


def test_side_effects():
    seen = []
    it = iter(IT.tee(fibonacci(), seen.append))
    want = [0, 1, 1, 2, 3, 5, 8, 13, 21, 34]
    for k in range(10):
        assert next(it) == want[k] and seen == want[: k + 1]


def test_sampling():
    rng = np.random.default_rng(0)
    it = rng.standard_normal(147)
    s = IT.sample(it, 10)
    assert len(s) == 15
    for k, x in enumerate(s):
        assert x == it[min(147, (k + 1) * 10) - 1]
    assert k == 14
    assert list(itertools.islice(IT.sample(fibonacci(), 3), 3)) == [1, 5, 21]
    with pytest.raises(ValueError):
        IT.sample(it, 0)


def test_timing():
    it = np.arange(5.0)
    timed = IT.stopwatch(it)
    assert len(timed) == 5
    for k, (t, x) in enumerate(timed):
        assert x == it[k] and t >= k * 2e7 * 0.9
        time.sleep(0.02)
