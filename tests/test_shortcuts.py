"""The two arithmetic shortcuts of round 2 that are argued rather than restated from the reference, checked on the CPU with
the same IEEE operations (numpy float32 / Python double; no GPU, no oracle):

  * csrc/traverse.cuh: trianglePlane rejects |a| < 4e-9 before t = a / dn.  Claim: with |dn| >= 0.00001f (the check in
    front of it, bvh.cpp:185) the quotient is below 0.0005 in magnitude whatever it is, so bvh.cpp:189 rejects it anyway.
  * csrc/wavefront.cu: powLobe evaluates cos^Ns for a whole Ns by squaring in double.  Claim: relative error below
    Ns * 2^-53 against the exact power, so the float it is rounded to equals (float)pow(x, Ns) except within that
    distance of a rounding boundary."""
import math

import numpy as np

T_MIN = np.float32(0.0005)        # bvh.cpp:189
DN_MIN = np.float32(0.00001)      # bvh.cpp:185
A_CUT = np.float32(4.0e-9)        # trianglePlane


def test_tiny_dividend_is_always_rejected():
    # the largest dividends that take the shortcut against the smallest divisors that reach it
    a_max = np.nextafter(A_CUT, np.float32(0))
    for a in (a_max, -a_max, np.float32(0.0), np.float32(-0.0), np.float32(1e-45), np.float32(1.17e-38)):
        for dn in (DN_MIN, -DN_MIN, np.nextafter(DN_MIN, np.float32(1)), np.float32(1.0), np.float32(3e38), np.float32(np.inf)):
            t = np.float32(a) / np.float32(dn)
            assert t < T_MIN, (a, dn, t)
    rng = np.random.default_rng(7)
    a = (rng.random(1 << 20, dtype=np.float32) * a_max).astype(np.float32) * rng.choice(np.float32([-1, 1]), 1 << 20)
    dn = np.exp(rng.uniform(math.log(1e-5), math.log(1e30), 1 << 20)).astype(np.float32)
    dn = np.maximum(dn, DN_MIN) * rng.choice(np.float32([-1, 1]), 1 << 20)
    assert np.all((a / dn) < T_MIN)
    # and the cut is not vacuous the other way: just above it a quotient CAN reach the threshold's neighbourhood
    assert np.float32(6e-9) / DN_MIN > T_MIN


def _pow_by_squaring(x, n):  # powLobe, operation for operation (Python floats are IEEE doubles, products are not fused)
    r = x if (n & 1) else 1.0
    b = x
    k = n >> 1
    while k:
        b = b * b
        if k & 1:
            r = r * b
        k >>= 1
    return r


def test_power_by_squaring_matches_pow():
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 50, 200, 250, 500, 1000, 4096):  # the cg22 materials use 1, 50, 200, 250, 500, 1000
        lo = 2.0 ** (-110.0 / n)  # below this the kernel takes the power as 0 without evaluating it
        xs = np.concatenate([rng.uniform(lo, 1.0, 20000), [lo, 1.0, np.nextafter(1.0, 0.0)]])
        differ = 0
        for x in xs:
            got, ref = _pow_by_squaring(float(x), n), math.pow(float(x), n)
            assert abs(got - ref) <= n * 2.0 ** -52 * ref, (n, x, got, ref)
            differ += np.float32(got) != np.float32(ref)
        assert differ <= 2, (n, differ)  # a float rounding boundary within 1e-13 of the value: about one case in 1e6
