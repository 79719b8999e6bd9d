"""The FP64 bicubic kernel evaluates Catmull-Rom in 19 operations where the reference spends 23 (csrc/exact.cuh,
catmull_rom_exact_h / catmull_rom_coef + catmull_rom_eval).  This pins the identities behind that form on the CPU: the
literal expression of GridH.cpp:215-217, rounded after every operation in C++ evaluation order, against the 19-operation
form with exactly rounded fused multiply-adds (emulated with rationals), bit for bit."""
import random
import struct
from fractions import Fraction


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))          # one rounding, like __fma_rn


def _bits(x):
    return struct.pack("<d", x)


def literal(p0, p1, p2, p3, t):                                     # GridH.cpp:215-217
    return 0.5 * (2 * p1 + (-p0 + p2) * t + (2 * p0 - 5 * p1 + 4 * p2 - p3) * t * t + (-p0 + 3 * p1 - 3 * p2 + p3) * t * t * t)


def kernel_form(p0, p1, p2, p3, t):                                 # exact.cuh: catmull_rom_exact_h with th = t / 2
    th = 0.5 * t
    lin = (p2 - p0) * th
    qc = _fma(4.0, p2, _fma(2.0, p0, -(5.0 * p1))) - p3
    quad = (qc * t) * th
    cc = ((3.0 * p1 - p0) - 3.0 * p2) + p3
    cub = ((cc * t) * t) * th
    return ((p1 + lin) + quad) + cub


def test_19_operation_catmull_rom_equals_the_reference_expression():
    rng = random.Random(20261018)
    ts = [0.0, 1e-13, 2.0 ** -40, 0.25, 0.5, 0.75, 1.0 - 1e-13, 1.0 - 2.0 ** -53]
    n = 0
    for k in range(6000):
        scale = rng.choice([1.0, 1e-3, 11000.0, 1e6])
        p = [rng.uniform(-1.0, 0.2) * scale for _ in range(4)]
        if k % 7 == 0:
            p[rng.randrange(4)] = rng.choice([0.0, -0.0])
        if k % 11 == 0:
            p[1] = p[2]                                             # flat stretch: cancellations to exact zero
        if k % 13 == 0:
            p[0] = p[2]
        for t in ts + [rng.random(), rng.random() * 1e-9]:
            want, got = literal(*p, t), kernel_form(*p, t)
            assert _bits(want) == _bits(got), (p, t, want, got)
            n += 1
    assert n == 6000 * 10
