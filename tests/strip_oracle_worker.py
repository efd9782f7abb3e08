"""CPU, gloo, world_size >= 2: the strip partition (clothgpu.strips.StripPlan) and the edge-row exchange plan, exercised
with the ORACLE as the per-strip solver (test infrastructure; the product exchanges inside its CUDA kernel).  Every rank
steps its strip, sends the edge rows its neighbours keep as ghosts (torch.distributed send/recv over gloo) and rank 0
checks the gathered strips against the whole-cloth oracle, bit for bit."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "cloth-physics_b200", "python"), os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def exchange(dist, torch, plan, rank, ghost, o, nx):
    """Send my owned edge rows to the neighbours that hold them as ghosts, receive theirs into my ghost rows."""
    up, dn = plan.neighbours(rank)
    r0, n = plan.rows(rank)
    held_lo, held_hi = plan.held(rank, ghost)
    reqs, recv = [], []
    # (neighbour, first of my rows it keeps, first of its rows I keep, how many)
    for nb, send_lo, recv_lo, rows in ((up, r0, held_lo, r0 - held_lo), (dn, r0 + n - (held_hi - r0 - n), r0 + n, held_hi - r0 - n)):
        if nb is None:
            continue
        x, y, px, py, fl = o.read_rows(send_lo, rows)
        out = torch.from_numpy(np.concatenate([x, y, px, py, fl.astype(np.float64)]))
        inn = torch.empty(5 * rows * nx, dtype=torch.float64)
        reqs += [dist.isend(out, nb), dist.irecv(inn, nb)]
        recv.append((recv_lo, rows, inn))
    for r in reqs:
        r.wait()
    for lo, rows, t in recv:
        a = t.numpy().reshape(5, rows * nx)
        o.write_rows(lo, rows, a[0], a[1], a[2], a[3], a[4].astype(np.uint8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=64)
    ap.add_argument("--ny", type=int, default=60)
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--frames", type=int, default=5)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    import pyoracle
    from clothgpu import _abi
    from clothgpu.strips import StripPlan

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    K = a.iterations
    ghost = 2 * K
    plan = StripPlan(a.ny, world, ghost=ghost)
    r0, n = plan.rows(rank)
    o = pyoracle.OracleCloth(_abi.grid_desc(a.nx, a.ny, pin_rule=_abi.PIN_REFERENCE, max_iterations=K,
                                            strip_row0=r0, strip_rows=n), pyoracle.COLOUR)
    assert (o.held_row0, o.held_row0 + o.held_rows) == plan.held(rank, ghost)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 128 + 6 * a.nx + 64, 164 + 6 * a.ny + 300
    f.buttons = _abi.MOUSE_DRAGGING | _abi.MOUSE_LEFT
    f.tear_length = 9.0
    f.mouse_force = 3.0
    whole = pyoracle.OracleCloth(_abi.grid_desc(a.nx, a.ny, pin_rule=_abi.PIN_REFERENCE, max_iterations=K),
                                 pyoracle.COLOUR) if rank == 0 else None
    ok = True
    for t in range(a.frames):
        b = plan.bounds[1]
        f.mouse_px, f.mouse_py = f.mouse_x, f.mouse_y
        f.mouse_x, f.mouse_y = 128.0 + 3 * a.nx + 7 * t, 164.0 + 6 * b - 5 + 4 * t
        o.step(f)
        exchange(dist, torch, plan, rank, ghost, o, a.nx)
        mine = o.read_rows(r0, n)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            whole.step(f)
            got = [np.concatenate([p[k] for p in parts]) for k in range(5)]
            for name, g_, r_ in zip(("x", "y", "px", "py", "flags"), got, whole.read_state()):
                if not np.array_equal(g_, r_):
                    print(f"frame {t}: plane {name} differs", flush=True)
                    ok = False
    if rank == 0 and ok:
        print(f"ORACLE-STRIPS-OK world={world} {a.nx}x{a.ny} K={K}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
