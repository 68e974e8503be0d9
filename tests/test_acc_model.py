"""Host-side model of the exact integer accumulators (csrc/srk_common.cuh acc_add / acc_read, include/srk.h "exact
sums"): the same digit arithmetic in Python integers, checked for the properties the design relies on - the split of a
float32 into radix-2^40 digits is exact, at most two digits are non-zero, the total of many partial sums is the exact
sum whatever the order, 64-bit counters cannot overflow with 148 contributors, non-finite values raise the flag.
(The device implementation itself is exercised by the -m gpu tests.)"""
import math
import random
import struct
from fractions import Fraction

import pytest

LIMBS = 5
UNIT_EXP = [-94 + 40 * k for k in range(LIMBS)]          # 2^-94, 2^-54, 2^-14, 2^26, 2^66
MASK64 = (1 << 64) - 1


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def acc_add(acc, flag, p):
    """acc: list of LIMBS unsigned 64-bit counters (Python ints mod 2^64); p: a float32 value."""
    if not math.isfinite(p) or abs(p) >= 2.0 ** 120:     # NaN / Inf / a partial sum that 148 contributors could overflow with
        flag[0] += 1
        return
    r = Fraction(p)                                       # the device works in float64: every step below is exact there
    for k in reversed(range(LIMBS)):
        unit = Fraction(2) ** UNIT_EXP[k]
        d = int(r / unit)                                 # truncation toward zero (cvt.rzi)
        r -= d * unit
        assert abs(d) < (1 << 62 if k == LIMBS - 1 else 1 << 40)
        acc[k] = (acc[k] + d) & MASK64                    # red.global.add.u64: two's complement
    return r                                              # what lies below 2^-94 is dropped


def acc_read(acc, flag):
    if flag[0]:
        return float("nan")
    total = Fraction(0)
    for k in range(LIMBS):
        d = acc[k] - (1 << 64) if acc[k] >> 63 else acc[k]
        total += d * Fraction(2) ** UNIT_EXP[k]
    return total


def test_digit_split_is_exact_and_touches_at_most_two_limbs():
    rng = random.Random(1)
    for _ in range(2000):
        p = f32(rng.uniform(-1, 1) * 2.0 ** rng.randint(-60, 100))
        acc, flag = [0] * LIMBS, [0]
        rest = acc_add(acc, flag, p)
        assert rest == 0 or abs(rest) < Fraction(2) ** -94
        assert sum(1 for d in acc if d) <= 2
        assert acc_read(acc, flag) + rest == Fraction(p)


def test_total_is_exact_and_order_independent():
    rng = random.Random(2)
    parts = [f32(rng.gauss(0, 1) * 10.0 ** rng.randint(-6, 9)) for _ in range(148)]
    want = sum(Fraction(p) for p in parts)
    totals = []
    for trial in range(4):
        rng.shuffle(parts)
        acc, flag = [0] * LIMBS, [0]
        for p in parts:
            acc_add(acc, flag, p)
        totals.append((tuple(acc), acc_read(acc, flag)))
    assert all(t == totals[0] for t in totals)            # the counters themselves do not depend on the order
    assert totals[0][1] == want                           # ... and hold the exact sum (nothing here is below 2^-94)


def test_counters_cannot_overflow_with_148_contributors():
    big = f32(2.0 ** 120 * (1 - 2.0 ** -24))              # the largest partial sum that is not flagged
    acc, flag = [0] * LIMBS, [0]
    for _ in range(512):
        acc_add(acc, flag, -big)
    assert acc_read(acc, flag) == 512 * Fraction(-big)
    acc_add(acc, flag, f32(2.0 ** 120))                   # at the bound: flagged, the total reads as NaN
    assert math.isnan(acc_read(acc, flag))
    lo = f32(2.0 ** 25 * (1 - 2.0 ** -24))                # largest value whose leading digit lands in a 40-bit limb
    acc, flag = [0] * LIMBS, [0]
    for _ in range(148):
        acc_add(acc, flag, lo)
    assert acc_read(acc, flag) == 148 * Fraction(lo)


@pytest.mark.parametrize("bad", [float("inf"), float("-inf"), float("nan")])
def test_non_finite_values_raise_the_flag(bad):
    acc, flag = [0] * LIMBS, [0]
    acc_add(acc, flag, 1.0)
    acc_add(acc, flag, bad)
    assert math.isnan(acc_read(acc, flag))
