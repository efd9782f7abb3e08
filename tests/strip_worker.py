"""Run under torch.distributed.run by tests/test_gpu_parity.py: every rank owns one row strip of a cloth on its GPU
(or all ranks share cuda:0 with --same-device), the strips exchange edge rows inside the fused kernel through IPC-mapped
peer memory, and rank 0 checks the gathered result against the whole cloth stepped on its own GPU, bit for bit."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "cloth-physics_b200", "python"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--iterations", type=int, default=1)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--schedule", default="auto", choices=["auto", "fused", "rowstream", "wave"])
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    import clothgpu
    from clothgpu import _abi
    from clothgpu.strips import StripCloth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = 0 if a.same_device else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    if a.backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")

    K = a.iterations
    rng = np.random.default_rng(1234)
    n = a.nx * a.ny
    gx, gy = np.meshgrid(np.arange(a.nx), np.arange(a.ny))
    x = (128.0 + 6.0 * gx + rng.uniform(-2.5, 2.5, gx.shape)).ravel()
    y = (164.0 + 6.0 * gy + rng.uniform(-2.5, 2.5, gy.shape)).ravel()
    px = x + rng.uniform(-0.5, 0.5, n)
    py = y + rng.uniform(-0.5, 0.5, n)
    fl = np.full(n, _abi.FLAG_ACTIVE | _abi.FLAG_ALIVE_H | _abi.FLAG_ALIVE_V, np.uint8)
    fl[rng.random(n) < 0.01] &= ~np.uint8(_abi.FLAG_ACTIVE)
    fl[rng.random(n) < 0.01] &= ~np.uint8(_abi.FLAG_ALIVE_H)
    fl[rng.random(n) < 0.01] &= ~np.uint8(_abi.FLAG_ALIVE_V)
    fl[: a.nx] |= _abi.FLAG_PIN

    f = _abi.default_frame(K)
    f.win_w, f.win_h = 128 + 6 * a.nx + 128, 164 + 6 * a.ny * 2
    f.buttons = _abi.MOUSE_DRAGGING
    f.tear_length = 11.0

    sched = {"auto": _abi.SCHED_AUTO, "fused": _abi.SCHED_FUSED, "rowstream": _abi.SCHED_ROWSTREAM, "wave": _abi.SCHED_WAVE}[a.schedule]
    s = StripCloth(a.nx, a.ny, max_iterations=K, device=dev)
    s.cloth.set_schedule(sched)
    s.write_state(x, y, px, py, fl)
    whole = None
    if rank == 0:
        whole = clothgpu.Cloth(_abi.grid_desc(a.nx, a.ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K, device=dev))
        whole.set_schedule(sched)
        whole.write_state(x, y, px, py, fl)
    ok = True
    for t in range(a.frames):
        # drag across a strip boundary so tears and pushes straddle it
        b = s.plan.bounds[1] if world > 1 else a.ny // 2
        f.mouse_px, f.mouse_py = f.mouse_x, f.mouse_y
        f.mouse_x, f.mouse_y = 128.0 + 6 * (a.nx // 2) + 9 * t, 164.0 + 6 * b - 4 + 3 * t
        s.step(f)
        got = s.gather_state()
        if rank == 0:
            whole.step(f)
            ref = whole.read_state()
            for name, g_, r_ in zip(("x", "y", "px", "py", "flags"), got, ref):
                if not np.array_equal(g_, r_):
                    bad = np.flatnonzero(g_ != r_)
                    print(f"frame {t}: plane {name} differs at {bad.size} particles, first rows {np.unique(bad // a.nx)[:8]}", flush=True)
                    ok = False
    cs = s.counters()
    if rank == 0:
        cw = whole.counters()
        for k in ("live_sticks", "alive_sticks", "relaxations", "inactive", "pinned"):
            if cs[k] != cw[k]:
                print(f"counter {k}: strips {cs[k]} != whole {cw[k]}", flush=True)
                ok = False
        if ok:
            print(f"STRIPS-OK world={world} backend={a.backend} {a.nx}x{a.ny} K={K} frames={a.frames} "
                  f"relaxations={cs['relaxations']} torn_total_alive={cw['alive_sticks']}", flush=True)
        whole.close()
    s.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
