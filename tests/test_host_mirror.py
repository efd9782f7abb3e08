"""The compiled host-side mirror of the reference's Go surface (cloth-physics_b200/host/physics.hpp: physics::Cloth,
physics::Mouse, gui::Hud over the C ABI), driven like main.go drives package physics."""
import json
import os
import subprocess

import numpy as np
import pytest

from clothgpu import _abi
from clothgpu.mouse import Mouse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "cloth-physics_b200", "host")
EXE = os.path.join(HOST, "physics_selftest")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "cloth-physics_b200", "csrc"), "-s"], check=True)
    subprocess.run(["make", "-C", HOST, "-s"], check=True)


def test_host_mirror_builds_and_fails_loudly_without_a_gpu():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is checked on the CPU box")
    r = subprocess.run([EXE, "5"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def _script(frames):
    """The same session physics_selftest.cc plays, as clothgpu_frame blocks."""
    f32 = np.float32
    m = Mouse()
    m.SetScrollY(50.0)
    m.SetMaxScrollY(120.0)
    for t in range(frames):
        if t == 20:
            m.SetRightButton()
        elif t == 60:
            m.ResetForce(); m.ReleaseLeftButton(); m.ReleaseRightButton(); m.SetDragging(False); m.SetCtrlDown(False)
        elif t == 70:
            m.SetLeftButton()
            m.UpdatePosition(400.0, 230.0)
        if 20 <= t < 60:
            m.UpdatePosition(float(f32(300.0) + f32(6.0) * f32(t - 20)), 260.0)
        if t > 70:
            m.SetForce((t - 70) / 60.0 * 5)
            m.SetLeftButton()
            m.UpdatePosition(float(f32(400.0) + f32(21.0) * f32(t - 70)), float(f32(230.0) + (f32(40.0) if (t // 3) % 2 else f32(-40.0))))
            m.SetDragging(True)
        f = _abi.default_frame(1)
        m.fill(f)
        yield f


@pytest.mark.gpu
def test_host_mirror_session_matches_the_oracle(oracle):
    _build()
    frames = 120
    r = subprocess.run([EXE, str(frames)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    o = oracle.OracleCloth(_abi.default_desc(), oracle.COLOUR)
    for f in _script(frames):
        o.step(f)
    x, y, px, py, fl = o.read_state()
    nx, ny = o.nx, o.ny
    act = (fl & _abi.FLAG_ACTIVE) != 0
    hl = (fl & _abi.FLAG_HIGHLIGHT) != 0
    X, Y, A, H, F = (a.reshape(ny, nx) for a in (x, y, act, hl, fl))
    live_h = ((F[:, :-1] & _abi.FLAG_ALIVE_H) != 0) & A[:, :-1] & A[:, 1:]
    live_v = ((F[:-1] & _abi.FLAG_ALIVE_V) != 0) & A[:-1] & A[1:]
    assert got["frames"] == frames
    assert got["relaxations"] == o.relaxations
    assert got["live_sticks"] == got["segments"] == int(live_h.sum() + live_v.sum())
    assert got["inactive"] == int((~act).sum()) and got["inactive"] > 0
    assert got["highlighted"] == int(hl.sum())
    assert got["alive_sticks"] == int(((fl & _abi.FLAG_ALIVE_H) != 0).sum() + ((fl & _abi.FLAG_ALIVE_V) != 0).sum())
    assert got["alive_sticks"] < 12105       # the drag tore sticks
    assert got["highlighted_segments"] == int((live_h & H[:, :-1] & H[:, 1:]).sum() + (live_v & H[:-1] & H[1:]).sum())
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    seg_sum = ((f32(X[:, :-1]) + f32(Y[:, :-1]) + f32(X[:, 1:]) + f32(Y[:, 1:]))[live_h].sum()
               + (f32(X[:-1]) + f32(Y[:-1]) + f32(X[1:]) + f32(Y[1:]))[live_v].sum())
    assert abs(got["segment_sum"] - seg_sum) <= 1e-6 * abs(seg_sum)
