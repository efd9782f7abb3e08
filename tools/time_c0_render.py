#!/usr/bin/env python
"""Where the end-to-end frame of the GUI cloth goes (1 GPU): step only, step + export (no copy), step + read_quads into pinned memory."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cloth-physics_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import clothgpu  # noqa: E402
from clothgpu import _abi  # noqa: E402
from traces import c0_trace  # noqa: E402

frames = list(c0_trace(1000))
cloth = clothgpu.Cloth(_abi.default_desc())
cap = 2 * 171 * 36
quads = torch.empty((cap, 8), dtype=torch.float32, pin_memory=True)
hl = torch.empty((cap,), dtype=torch.uint8, pin_memory=True)


def run(kind):
    cloth.reset(128, 164)
    t0 = time.perf_counter()
    for f in frames:
        cloth.step(f)
        if kind == "step":
            cloth.sync()
        elif kind == "export":
            cloth.compact_segments()
            cloth.sync()
        elif kind == "count":
            cloth.compact_segments()
            cloth.count_segments()
        else:
            cloth.read_quads_into(quads.data_ptr(), hl.data_ptr(), cap)
    return (time.perf_counter() - t0) / len(frames) * 1e6


for kind in ("step", "export", "count", "quads"):
    run(kind)
    print(f"{kind:7s} {min(run(kind) for _ in range(3)):7.1f} us per frame")
