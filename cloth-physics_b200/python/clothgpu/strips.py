"""One cloth over several GPUs: row strips, one process per GPU.

`StripPlan` is the partition (pure host logic, also exercised on CPU by tests/test_strips_gloo.py); `StripCloth`
is one rank's strip on its GPU.  torch.distributed is used for the rendezvous only — an all_gather of the
small blobs `clothgpu_strip_export` produces.  After `clothgpu_strip_attach` the per-frame exchange is done by
the fused frame kernel itself: it stores the edge rows a neighbour keeps as ghosts straight into the neighbour's
memory over NVLink and hand-shakes with stream-ordered epoch flags (csrc/cloth_kernels.cuh).  There is no
data-path collective and no CPU path.
"""
import numpy as np

from . import _abi


class StripPlan:
    """Rows [bounds[r], bounds[r+1]) belong to rank r.  Every boundary is even, so the vertical sticks that straddle a
    boundary all have the same colour (V-odd) — the reason strips must start on an even row (clothgpu_create)."""

    def __init__(self, ny, world, ghost=0):
        if world < 1 or ny < 2 * world:
            raise ValueError(f"cannot cut {ny} rows into {world} strips of at least 2 rows")
        b = [2 * ((i * ny + world) // (2 * world)) for i in range(world)] + [ny]
        b[0] = 0
        self.ny, self.world, self.bounds = ny, world, b
        for r in range(world):
            n = b[r + 1] - b[r]
            if n < max(2, ghost):
                raise ValueError(f"strip {r} would own {n} rows, fewer than the {ghost} ghost rows its neighbours keep")

    def rows(self, rank):
        return self.bounds[rank], self.bounds[rank + 1] - self.bounds[rank]

    def held(self, rank, ghost):
        """Global rows [lo, hi) a strip holds: its own rows plus `ghost` rows of each neighbour."""
        r0, n = self.rows(rank)
        return max(0, r0 - ghost), min(self.ny, r0 + n + ghost)

    def neighbours(self, rank):
        return (rank - 1 if rank > 0 else None), (rank + 1 if rank + 1 < self.world else None)


def ghost_rows(max_iterations, sweeps_per_launch=8):
    """Ghost depth clothgpu_create gives a strip: 2 rows per sweep (csrc/clothgpu_api.cu)."""
    return 2 * max(1, max_iterations)


class StripCloth:
    """This rank's strip of one nx x ny cloth.  All ranks must call the same methods in the same order."""

    def __init__(self, nx, ny, *, max_iterations=1, pin_rule=_abi.PIN_TOP_ROW, precision=_abi.F64, device=None,
                 ensemble=1, group=None):
        import torch
        import torch.distributed as dist

        from . import Cloth

        self._dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device()
        self.plan = StripPlan(ny, self.world, ghost=2 * max_iterations if self.world > 1 else 0)
        row0, rows = self.plan.rows(self.rank)
        desc = _abi.grid_desc(nx, ny, pin_rule=pin_rule, precision=precision, ensemble=ensemble, device=device,
                              max_iterations=max_iterations, strip_row0=row0, strip_rows=rows)
        self.cloth = Cloth(desc)
        self.nx, self.ny, self.row0, self.rows = nx, ny, row0, rows
        if self.world > 1:
            blobs = [None] * self.world
            dist.all_gather_object(blobs, self.cloth.strip_export(), group=group)
            up, dn = self.plan.neighbours(self.rank)
            self.cloth.strip_attach(None if up is None else blobs[up], None if dn is None else blobs[dn])
            dist.barrier(group=group)  # nobody steps before every strip is attached

    def close(self):
        if getattr(self, "cloth", None) is not None:
            self.cloth.sync()
            if self.world > 1:
                self._dist.barrier(group=self.group)  # a neighbour may still be storing into my ghost rows
            self.cloth.close()
            self.cloth = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self, frame):
        self.cloth.step(frame)

    def step_n(self, frames):
        self.cloth.step_n(frames)

    def sync(self):
        self.cloth.sync()

    def write_state(self, x=None, y=None, px=None, py=None, flags=None):
        """Arrays of the WHOLE cloth (every rank passes the same); each rank keeps its rows, then the edge rows are
        pushed into the neighbours' ghost rows."""
        lo, hi = self.row0 * self.nx, (self.row0 + self.rows) * self.nx
        part = [None if a is None else np.ascontiguousarray(a).ravel()[lo:hi] for a in (x, y, px, py, flags)]
        self.cloth.write_state(*part)
        if self.world > 1:
            self.cloth.strip_push()

    def read_state(self):
        """This rank's owned rows."""
        return self.cloth.read_state()

    def gather_state(self):
        """The whole cloth on every rank (tests)."""
        mine = self.cloth.read_state()
        if self.world == 1:
            return mine
        parts = [None] * self.world
        self._dist.all_gather_object(parts, mine, group=self.group)
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(5))

    def counters(self):
        """Counters summed over the strips."""
        c = self.cloth.counters()
        if self.world == 1:
            return c
        allc = [None] * self.world
        self._dist.all_gather_object(allc, c, group=self.group)
        out = {k: sum(a[k] for a in allc) for k in c}
        out["frames"] = c["frames"]
        return out
