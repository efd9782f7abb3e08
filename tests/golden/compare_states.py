#!/usr/bin/env python
"""Diff two directories of CLSTATE1 snapshots (clothgpu/statefile.py) file by file, bit for bit:

    python tests/golden/make_golden.py --dump-states /tmp/oracle         # the oracle, all three modes
    go run -tags refharness ./harness -trace tests/golden/c0_trace.jsonl -checkpoints ... -out /tmp/ref/c0_seq_faithful
    python tests/golden/compare_states.py /tmp/ref /tmp/oracle

The day a Go toolchain is available this is the comparison that pins the oracle to the real reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "cloth-physics_b200", "python"))
from clothgpu import statefile  # noqa: E402


def compare(a_path, b_path):
    (ha, ca), (hb, cb) = statefile.read(a_path), statefile.read(b_path)
    out = []
    for k in ("nx", "ny", "row0", "rows", "ensemble"):
        if ha[k] != hb[k]:
            return [f"header {k}: {ha[k]} != {hb[k]}"]
    for e, (sa, sb) in enumerate(zip(ca, cb)):
        for name, pa, pb in zip(("x", "y", "px", "py", "flags"), sa, sb):
            if not np.array_equal(pa.view(np.uint64) if pa.dtype == np.float64 else pa, pb.view(np.uint64) if pb.dtype == np.float64 else pb):
                bad = np.flatnonzero(pa != pb)
                out.append(f"cloth {e} plane {name}: {bad.size} of {pa.size} differ, first {bad[:4].tolist()}: {pa[bad[:4]]} vs {pb[bad[:4]]}")
    return out


def main():
    a, b = sys.argv[1:3]
    names = sorted(set(os.listdir(a)) & set(os.listdir(b)))
    if not names:
        print("no common file names")
        return 2
    bad = 0
    for n in names:
        diffs = compare(os.path.join(a, n), os.path.join(b, n))
        print(("DIFF " if diffs else "same ") + n)
        for d in diffs:
            print("    " + d)
        bad += bool(diffs)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
