#!/usr/bin/env python
"""Generate tests/golden/c0_golden.json: state digests of the scripted default-cloth trace (BASELINE configs[0]:
171 x 36 particles, reference pin rule, 1000 frames, tests/traces.py::c0_trace) as computed by the CPU oracle.

PARITY UNPINNED: the reference cannot be built here (no Go toolchain), so these vectors come from the oracle itself.
They freeze its behaviour (compiler / flag / refactoring regressions) and give the GPU tests an oracle-independent
target at run time.  With a Go toolchain, host/go/harness regenerates them from the real reference.

    python tests/golden/make_golden.py                # rewrites c0_golden.json
    python tests/golden/make_golden.py --dump-trace   # rewrites c0_trace.jsonl: the scripted trace as JSON lines
                                                      # (clothgpu/trace.py), the input of host/go/harness/main.go
    python tests/golden/make_golden.py --dump-states DIR   # CLSTATE1 snapshots of the oracle (all three modes) at the
                                                      # checkpoints, to diff against the Go harness's dumps
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "cloth-physics_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CHECKPOINTS = (0, 3, 99, 140, 399, 449, 549, 599, 799, 800, 999)
SAMPLES = (0, 17, 170, 171, 3000, 4321, 6155)          # particle indices whose values are stored verbatim
MODES = {"seq_faithful": 0, "seq_masked": 1, "colour": 2}
ITERATIONS = {"seq_faithful": (1,), "seq_masked": (1,), "colour": (1, 3)}


def digest(state):
    out = {}
    for name, a in zip(("x", "y", "px", "py", "flags"), state):
        out[name] = hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return out


def run(pyoracle, _abi, c0_trace, mode, K):
    o = pyoracle.OracleCloth(_abi.default_desc(), MODES[mode])
    recs = {}
    for t, f in enumerate(c0_trace(1000, iterations=K)):
        o.step(f)
        if t in CHECKPOINTS:
            st = o.read_state()
            fl = st[4]
            recs[str(t)] = {
                "sha256": digest(st),
                "relaxations": o.relaxations,
                "slice_len": o.slice_len,          # len(cloth.constraints); only the sequential modes shrink the slice
                "alive_sticks": int((fl & _abi.FLAG_ALIVE_H != 0).sum() + (fl & _abi.FLAG_ALIVE_V != 0).sum()),
                "torn_last": o.torn_last,
                "pinned": int((fl & _abi.FLAG_PIN != 0).sum()),
                "inactive": int((fl & _abi.FLAG_ACTIVE == 0).sum()),
                "highlighted": int((fl & _abi.FLAG_HIGHLIGHT != 0).sum()),
                "samples": {str(i): [float(st[0][i]).hex(), float(st[1][i]).hex(), float(st[2][i]).hex(),
                                     float(st[3][i]).hex(), int(fl[i])] for i in SAMPLES},
            }
    return recs


def dump_trace(path):
    from clothgpu.trace import RecordingMouse, TraceRecorder
    from traces import c0_trace
    rec = TraceRecorder(RecordingMouse())
    c0_trace(1000, recorder=rec)
    rec.save(path)
    print("wrote", path, len(rec.records), "frames")


def dump_states(out_dir):
    import pyoracle
    from clothgpu import _abi, statefile
    from traces import c0_trace
    os.makedirs(out_dir, exist_ok=True)
    for mode, code in MODES.items():
        o = pyoracle.OracleCloth(_abi.default_desc(), code)
        for t, f in enumerate(c0_trace(1000)):
            o.step(f)
            if t in CHECKPOINTS:
                statefile.write(os.path.join(out_dir, f"c0_{mode}_{t:04d}.clstate"), 171, 36, [o.read_state()], frames=t + 1)
    print("wrote", out_dir)


def main():
    if "--dump-trace" in sys.argv:
        return dump_trace(os.path.join(HERE, "c0_trace.jsonl"))
    if "--dump-states" in sys.argv:
        return dump_states(sys.argv[sys.argv.index("--dump-states") + 1])
    import pyoracle
    from clothgpu import _abi
    from traces import c0_trace
    gold = {"what": "oracle digests of the c0 trace; see make_golden.py", "nx": 171, "ny": 36, "frames": 1000, "runs": {}}
    for mode, ks in ITERATIONS.items():
        for K in ks:
            gold["runs"][f"{mode}/K{K}"] = run(pyoracle, _abi, c0_trace, mode, K)
    with open(os.path.join(HERE, "c0_golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "c0_golden.json"))


if __name__ == "__main__":
    main()
