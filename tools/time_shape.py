#!/usr/bin/env python
"""Time the frame kernel on an arbitrary cloth shape (1 GPU): tools/time_shape.py NX NY K [frames] [schedule]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cloth-physics_b200", "python"))
import clothgpu  # noqa: E402
from clothgpu import _abi  # noqa: E402

nx, ny, K = (int(a) for a in sys.argv[1:4])
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 50
prec = _abi.F32 if os.environ.get("CLOTH_F32") else _abi.F64   # CLOTH_F32=1: the binary32 mode
cloth = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K, precision=prec))
if len(sys.argv) > 5:
    cloth.set_schedule(getattr(_abi, "SCHED_" + sys.argv[5].upper()))
f = _abi.default_frame(K)
f.win_w, f.win_h = 128 + 6 * nx + 128, 164 + 6 * ny * 4
if os.environ.get("CLOTH_DRAG"):      # the DRAG instantiation with the mouse far away: nothing tears, the tear test runs
    f.buttons = _abi.MOUSE_LEFT | _abi.MOUSE_DRAGGING
    f.mouse_x = f.mouse_y = f.mouse_px = f.mouse_py = -1.0e6
stream = torch.cuda.Stream()
cloth.set_stream(stream.cuda_stream)
arr = clothgpu.frame_array([f] * frames)
cloth.step_n(clothgpu.frame_array([f] * 5))
cloth.sync()
with torch.cuda.stream(stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    cloth.step_n(arr)
    e1.record(stream)
e1.synchronize()
ms = e0.elapsed_time(e1) / frames
print(f"{nx}x{ny} K={K}: {ms:.4f} ms/frame, {65.1 * nx * ny / ms / 1e6:.0f} GB/s compulsory, "
      f"{K * (2 * nx * ny - nx - ny) / ms / 1e6:.1f} G relax/s")
