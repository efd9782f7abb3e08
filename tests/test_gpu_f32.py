"""The optional binary32 mode (BASELINE.json: "stored as float64 for parity with an optional fp32 mode") against the oracle
run in binary32 (oracle/cloth_oracle.c instantiates its implementation for float as well).

First run on the round-1 driver box (GPUTEST_r01.json: XPASS on every case); the xfail marker it carried until then is
gone, so a binary32 mismatch now fails the suite.
"""
import numpy as np
import pytest

import clothgpu
from clothgpu import _abi
from clothgpu._abi import MOUSE_DRAGGING, SCHED_FUSED, SCHED_ROWSTREAM, SCHED_STREAMING, SCHED_WAVE

from test_gpu_parity import assert_same_state, random_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,ny,K,schedule", [(171, 36, 1, SCHED_ROWSTREAM), (171, 36, 1, SCHED_FUSED), (130, 131, 3, SCHED_FUSED),
                                              (700, 219, 1, SCHED_ROWSTREAM), (1037, 90, 1, SCHED_ROWSTREAM), (126, 300, 1, SCHED_ROWSTREAM),
                                              (3, 40, 1, SCHED_ROWSTREAM), (129, 64, 1, SCHED_ROWSTREAM),
                                              (300, 219, 8, SCHED_WAVE), (129, 33, 2, SCHED_STREAMING)])
def test_f32_mode_matches_the_binary32_oracle(oracle, nx, ny, K, schedule):
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=K, precision=_abi.F32)
    g = clothgpu.Cloth(desc)
    g.set_schedule(schedule)
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    st = random_state(nx, ny, seed=nx + ny + K)
    st = tuple(a.astype(np.float32).astype(np.float64) if a.dtype == np.float64 else a for a in st)
    g.write_state(*st)
    o.write_state(*st)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 128 + 6 * nx + 50, 164 + 6 * ny + 20
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 9.5
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 128.0 + 3 * nx, 164.0 + 3 * ny, 120.0 + 3 * nx, 170.0 + 3 * ny
    for t in range(3):
        g.step(f)
        o.step(f)
        assert_same_state(g.read_state(), o.read_state(), f"f32 {nx}x{ny} K={K} sched={schedule} frame {t}")
    assert g.counters()["relaxations"] == o.relaxations
