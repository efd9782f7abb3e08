import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cloth-physics_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import clothgpu, pyoracle
from clothgpu import _abi
from test_gpu_parity import random_state

nx = ny = 128
E = 5
for sched in (_abi.SCHED_STREAMING, _abi.SCHED_FUSED):
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, ensemble=E)
    g = clothgpu.Cloth(desc); g.set_schedule(sched)
    one = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW)
    oracles = [pyoracle.OracleCloth(one, pyoracle.COLOUR) for _ in range(E)]
    for e in range(E):
        st = random_state(nx, ny, seed=100 + e, pins=0.0)
        g.write_state(*st, cloth=e)
        oracles[e].write_state(*st)
    frames = []
    for e in range(E):
        f = _abi.default_frame(8)
        f.win_w, f.win_h = 2000, 2000
        f.buttons = _abi.MOUSE_DRAGGING if e % 2 else 0
        f.mouse_x, f.mouse_y = 128.0 + 40 * e, 300.0
        f.mouse_px, f.mouse_py = f.mouse_x - 3, f.mouse_y + 2
        f.tear_length = 10.0
        frames.append(f)
    for t in range(4):
        g.step_each(frames)
        for e in range(E):
            oracles[e].step(frames[e])
            gs, os_ = g.read_state(cloth=e), oracles[e].read_state()
            for name, a, b in zip("x y px py fl".split(), gs, os_):
                bad = np.flatnonzero(a != b)
                if bad.size:
                    if name != "fl":
                        ulp = np.abs(a[bad] - b[bad]) / np.spacing(np.abs(b[bad]))
                        print("sched", sched, "frame", t, "cloth", e, name, bad.size, "first", bad[:6], "ulps", ulp[:6], "max", ulp.max())
                    else:
                        print("sched", sched, "frame", t, "cloth", e, name, bad.size, bad[:6], a[bad[:6]], b[bad[:6]])
    print("sched", sched, "done; torn", g.counters()["torn_last"], sum(o.torn_last for o in oracles), "relax", g.counters()["relaxations"], sum(o.relaxations for o in oracles))
