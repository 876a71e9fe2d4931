"""Leadership.java:105 computes `Math.round(Math.log(Math.E + recentRejection))` in double arithmetic.
The kernel uses an integer threshold table instead; this checks the table against libm at every
threshold, around it, and exhaustively for r < 2^17."""
import ctypes as C
import math

from rafting_b200 import engine


def jvm(r):
    return math.floor(math.log(math.e + r) + 0.5)      # Math.round(double) == floor(x + 0.5)


def test_table_matches_double_math():
    L = engine.lib()
    L.rafting_backoff_step.restype = C.c_int64
    L.rafting_backoff_step.argtypes = [C.c_int32]
    for r in range(0, 1 << 17):
        assert L.rafting_backoff_step(r) == jvm(r), r
    # every threshold e^(k+0.5) - e up to int32 max, +-3 around it
    k = 1
    while True:
        t = math.exp(k + 0.5) - math.e
        if t > 2**31 - 1:
            break
        for r in range(max(0, int(t) - 3), min(2**31 - 1, int(t) + 4)):
            assert L.rafting_backoff_step(r) == jvm(r), (k, r)
        k += 1
    assert k >= 21
    for r in (2**31 - 1, 2**30, 10**9):
        assert L.rafting_backoff_step(r) == jvm(r)
