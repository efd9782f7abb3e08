#!/usr/bin/env python
"""Hard cases of the reciprocal step of StickMath<double> (csrc/cloth_fused.cuh).

r = fma(y1, e2, y1), e2 = fma(y1, -dist, 1): one Newton step from y1 ~ 1/sqrt(s).  With |y1*dist - 1| <= 2^-52 + 2^-60
(half an ulp of y1's own rounding, the 2^-53 relative rounding error of dist, the cubic's truncation error), e2 is exact
(a multiple of 2^-105 below 2^-52) and the value rounded by the last fma is (1/dist)(1 - eps^2), eps^2 <= 2^-104 * 1.01.
r differs from RN(1/dist) only if a rounding midpoint m = M*2^-54 (M odd) of the binade (1/2, 1] lies within that distance
of 1/dist, dist = B*2^-52 in [1, 2):  |2^106 - B*M| <= 2^106 * eps^2 ~ 4.  This script enumerates every such (B, M) for
|k| = |2^106 - B*M| <= KMAX (a generous margin) by factoring 2^106 - k with GNU factor, and prints the dist mantissas a
device test must cover (tests/test_gpu_stickmath.py feeds them to clothgpu_selftest_stickmath for every exponent parity
and every rest length).
"""
import itertools
import json
import subprocess
import sys

KMAX = 16


def divisors(n):
    out = subprocess.run(["factor", str(n)], capture_output=True, text=True, check=True).stdout
    primes = [int(p) for p in out.split(":")[1].split()]
    ds = {1}
    for p in primes:
        ds |= {d * p for d in ds}
    return sorted(ds)


def hard_cases(kmax=KMAX):
    found = []
    for k in itertools.chain(range(1, kmax + 1), range(-kmax, 0)):
        n = 2 ** 106 - k
        for b in divisors(n):
            if 2 ** 52 <= b < 2 ** 53:
                m = n // b
                if m % 2 == 1 and 2 ** 53 < m < 2 ** 54:
                    found.append({"k": k, "B": b, "M": m})
    return found


if __name__ == "__main__":
    cases = hard_cases()
    for c in cases:
        print(f"k={c['k']:+d}  B=0x{c['B']:014x}  (dist mantissa 1.{c['B'] - 2**52:013x})  M=0x{c['M']:014x}")
    if len(sys.argv) > 1:
        json.dump(cases, open(sys.argv[1], "w"), indent=1)
