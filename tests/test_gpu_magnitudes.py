"""The shared-seed sqrt / reciprocal / quotient expansion of the stick arithmetic (cloth_fused.cuh, StickMath<double>) on
stick lengths far from the usual few pixels: coordinates spread over 1 .. 2^30 (the window is an int32), i.e. squared
lengths between the rest length squared and ~2^61, with random mantissas — against the oracle's libm sqrt and `/`.

First run on the round-1 driver box (GPUTEST_r01.json: XPASS on all 15 cases); the xfail marker is gone.  The
exhaustive-style evidence for the expansion is tests/test_gpu_stickmath.py (2^34 random operands and the hard cases of
the reciprocal step on the device against sqrt.rn / div.rn).
"""
import numpy as np
import pytest

import clothgpu
from clothgpu import _abi
from clothgpu._abi import FLAG_ACTIVE, FLAG_ALIVE_H, FLAG_ALIVE_V, SCHED_FUSED, SCHED_STREAMING, SCHED_WAVE

from test_gpu_parity import assert_same_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("schedule,K", [(SCHED_STREAMING, 1), (SCHED_FUSED, 2), (SCHED_WAVE, 4)])
@pytest.mark.parametrize("log2_scale", [0, 8, 16, 24, 29])
def test_stick_arithmetic_on_wide_magnitudes(oracle, schedule, K, log2_scale):
    nx, ny = 96, 40
    desc = _abi.grid_desc(nx, ny, spacing=1, pin_rule=_abi.PIN_NONE, max_iterations=K)
    g = clothgpu.Cloth(desc)
    g.set_schedule(schedule)
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    rng = np.random.default_rng(log2_scale + 7 * K)
    scale = float(2 ** log2_scale)
    # neighbours anywhere from ~1 to ~scale apart, random mantissas; everything inside the window [0, 2^30]
    x = rng.uniform(0.0, scale, nx * ny) + np.repeat(np.arange(ny), nx) * 0.37
    y = rng.uniform(0.0, scale, nx * ny) + np.tile(np.arange(nx), ny) * 0.53
    px = x + rng.uniform(-0.25, 0.25, x.shape)
    py = y + rng.uniform(-0.25, 0.25, y.shape)
    fl = np.full(nx * ny, FLAG_ACTIVE | FLAG_ALIVE_H | FLAG_ALIVE_V, np.uint8)
    fl.reshape(ny, nx)[:, -1] &= ~np.uint8(FLAG_ALIVE_H)
    fl.reshape(ny, nx)[-1, :] &= ~np.uint8(FLAG_ALIVE_V)
    g.write_state(x, y, px, py, fl)
    o.write_state(x, y, px, py, fl)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 2 ** 30, 2 ** 30
    f.slider[1] = 0.0          # no gravity: the lengths stay spread out
    for t in range(4):
        g.step(f)
        o.step(f)
        assert_same_state(g.read_state(), o.read_state(), f"scale 2^{log2_scale} K={K} sched={schedule} frame {t}")
