"""Multi-process host logic on CPU: gloo backend, world_size 2 and 3 (no GPU needed)."""
import os
import socket
import subprocess
import sys

import pytest

from clothgpu.strips import StripPlan

HERE = os.path.dirname(os.path.abspath(__file__))


def test_strip_plan_partitions_rows_on_even_boundaries():
    for ny, world in ((16384, 8), (16384, 2), (1024, 4), (211 // 6 + 1, 2), (37, 3), (1000, 7)):
        p = StripPlan(ny, world)
        assert p.bounds[0] == 0 and p.bounds[-1] == ny
        assert all(b % 2 == 0 for b in p.bounds[:-1])
        sizes = [p.rows(r)[1] for r in range(world)]
        assert sum(sizes) == ny and min(sizes) >= 2 and max(sizes) - min(sizes) <= 3
        for r in range(world):
            lo, hi = p.held(r, 4)
            r0, n = p.rows(r)
            assert lo == max(0, r0 - 4) and hi == min(ny, r0 + n + 4)
    assert StripPlan(16384, 8).rows(3) == (6144, 2048)
    with pytest.raises(ValueError):
        StripPlan(40, 4, ghost=16)       # 10-row strips cannot serve 16 ghost rows
    with pytest.raises(ValueError):
        StripPlan(3, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,K", [(2, 1), (2, 3), (3, 2)])
def test_oracle_strips_with_gloo_exchange_equal_the_whole_cloth(world, K):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "strip_oracle_worker.py"), "--nx", "70", "--ny", "66", "--iterations", str(K), "--frames", "6"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ORACLE-STRIPS-OK" in r.stdout
