"""The binary64 stick arithmetic of every fast kernel — sqrt(s), (L - d)/d and L/d from one reciprocal-square-root seed
(csrc/cloth_fused.cuh, StickMath<double>) — against the device's IEEE sqrt.rn.f64 / div.rn.f64 (what Go's SQRTSD / DIVSD
compute in physics/constraint.go:28,41-42), through the C ABI's self-test hook:

  * 2^34 pseudo-random stretched sticks (rest length 1..64, squared length up to 2^61, random mantissas), and
  * the hard cases of the reciprocal step: every stick length whose reciprocal lies within 16 * 2^-106 (relative) of a
    rounding midpoint (tools/stickmath_hard_cases.py enumerates them by factoring 2^106 - k), at every binary exponent
    the window allows and a few far beyond, each with its neighbouring squared lengths and every rest length 1..64.

Zero mismatches is the contract: one differing bit would break bit-exactness against the oracle somewhere.
"""
import math
import os
import sys

import numpy as np
import pytest

import clothgpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from stickmath_hard_cases import hard_cases  # noqa: E402

pytestmark = pytest.mark.gpu


def squared_lengths_rounding_to(B, e):
    """Doubles s around (B * 2^(e-52))^2: the ones whose square root rounds to that length, and their neighbours."""
    d = math.ldexp(float(B), e - 52)     # exact: B has 53 bits
    s0 = d * d                           # RN(d^2)
    out = [s0]
    lo = hi = s0
    for _ in range(3):
        lo, hi = math.nextafter(lo, 0.0), math.nextafter(hi, math.inf)
        out += [lo, hi]
    return out


def test_random_operands_match_ieee_sqrt_and_div():
    rep = clothgpu.selftest_stickmath(samples=1 << 34, seed=20261020)
    assert rep["tested"] == 1 << 34
    assert rep["mismatches"] == 0, rep


def test_hard_cases_of_the_reciprocal_step():
    cases = hard_cases()
    assert any(c["B"] == 2 ** 53 - 1 for c in cases), "the all-ones mantissa is the classical hard case"
    s = []
    for c in cases:
        for e in list(range(0, 31)) + [40, 77, 200, 401]:
            s += squared_lengths_rounding_to(c["B"], e)
    s = np.array([v for v in s if 1.0 <= v < 2.0 ** 899], np.float64)
    # every listed length really is hit: sqrt of the centre value rounds back to B * 2^(e-52)
    d = math.ldexp(float(cases[0]["B"]), 3 - 52)
    assert math.sqrt(d * d) == d
    rep = clothgpu.selftest_stickmath(s_list=s)
    assert rep["tested"] > 30 * len(s)
    assert rep["mismatches"] == 0, rep


def test_binary32_expansion_is_exact_on_every_operand():
    """The binary32 mode's expansion, exhaustively: every float s in [1, 2^63) with every rest length 1..64 that leaves the
    stick stretched, against sqrt.rn.f32 / div.rn.f32 — a proof by enumeration, not a sample."""
    rep = clothgpu.selftest_stickmath_f32()
    assert rep["tested"] > 2.8e10, rep
    assert rep["mismatches"] == 0, rep


def test_selftest_rejects_operands_outside_its_range():
    with pytest.raises(clothgpu.ClothGpuError):
        clothgpu.selftest_stickmath(s_list=np.array([0.5]))
