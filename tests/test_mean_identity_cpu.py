"""bilinear_fill_kernel forms the reference's NaN-corner mean (GridH.cpp:10-18, :186-198) without branches (csrc/fill.cu,
mean_valid4_flat): +0.0 stands in for a missing corner, and the division by the count is a product with a tabulated
reciprocal plus one Markstein step.  Both identities, bit for bit, on the CPU."""
import math
import random
import struct
from fractions import Fraction

INV = [math.nan, 1.0, 0.5, 0.33333333333333331, 0.25]


def _fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _bits(x):
    return struct.pack("<d", x)


def reference_mean(vals):                                           # fallbackAverage: skip the NaNs, sum / count
    s, n = 0.0, 0
    for v in vals:
        if not math.isnan(v):
            s += v
            n += 1
    return s / n if n else math.nan


def kernel_mean(vals):
    n = sum(0 if math.isnan(v) else 1 for v in vals)
    s = 0.0
    for v in vals:
        s = s + (0.0 if math.isnan(v) else v)
    if n == 0:
        return math.nan
    inv, cnt = INV[n], float(n)
    mag = abs(s)
    if not (1e-280 < mag < 1e300) and s != 0.0:                     # the kernel's guard: the division itself (out of line)
        return s / cnt
    q = s * inv
    r = _fma(-cnt, q, s)
    return _fma(r, inv, q)


def test_branch_free_corner_mean_equals_the_reference_mean():
    rng = random.Random(7)
    import numpy as np
    for k in range(40000):
        scale = rng.choice([1.0, 1e-4, 11000.0, 3.0, 1e-290, 1e305])
        as_f32 = k % 2 == 1 and 1e-30 < scale < 1e30                 # FP32 grids: values exactly representable in float
        vals = [float(np.float32(rng.uniform(-1.0, 0.1) * scale)) if as_f32 else rng.uniform(-1.0, 0.1) * scale for _ in range(4)]
        for j in range(4):
            if rng.random() < 0.45:
                vals[j] = math.nan
        if k % 17 == 0:
            vals[rng.randrange(4)] = rng.choice([0.0, -0.0])
        if k % 19 == 0:                                             # multiples of three and their neighbours
            m = float(rng.randrange(1, 1 << 50))
            vals = [m, m, m + rng.choice([0.0, 1.0, -1.0]), math.nan]
        want, got = reference_mean(vals), kernel_mean(vals)
        assert (math.isnan(want) and math.isnan(got)) or _bits(want) == _bits(got), (vals, want, got)
