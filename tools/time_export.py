#!/usr/bin/env python
"""Time the live-stick export kernel alone on an nx x ny cloth (1 GPU): tools/time_export.py NX NY [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cloth-physics_b200", "python"))
import clothgpu  # noqa: E402
from clothgpu import _abi  # noqa: E402

nx, ny = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cloth = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=1))
stream = torch.cuda.Stream()
cloth.set_stream(stream.cuda_stream)
f = _abi.default_frame(1)
f.win_w, f.win_h = 128 + 6 * nx + 128, 164 + 6 * ny * 4
cloth.step_n(clothgpu.frame_array([f] * 3))
for _ in range(3):
    cloth.compact_segments()
cloth.sync()
with torch.cuda.stream(stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        cloth.compact_segments()
    e1.record(stream)
e1.synchronize()
ms = e0.elapsed_time(e1) / reps
n = cloth.count_segments()
print(f"{nx}x{ny}: export {ms * 1e3:.1f} us, {n} sticks, {(16.0 * nx * ny + 17.0 * n + nx * ny) / ms / 1e6:.0f} GB/s")
