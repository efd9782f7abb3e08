"""A model check of the wavefront kernel's synchronisation (cloth-physics_b200/csrc/cloth_wave.cuh) — runs on the CPU.

compute-sanitizer's racecheck is not available on the GPU pool, so the happens-before structure of the kernel is checked
here on a model: the events of a chunk (entry work E(s), phase alpha A(j, s) and phase beta B(j, s) of every sweep stage
j and step s), the ring slots each of them reads or writes, and the ordering edges the kernel's barriers give:

  * alpha -> beta between the two worker groups that share their stages (128 threads; local stages {0, 2} / {1, 3} of a set)
  * end of step within a set (set X's barrier includes the entry warps)
  * "X has finished step s" -> set Y may start step s + 1          (bar.arrive / bar.sync, ids 4..7)
  * "Y has finished step s" -> the entry warps may start step s + 2 (ids 8..11)

Two properties must hold for every sweep count the kernel supports:
  1. every value an event consumes was produced by an event that happens before it (the data dependencies of the
     colour-ordered sweeps), and
  2. two events that are NOT ordered never touch the same ring slot (every access is a read-modify-write).
The GPU tests check the results bit for bit; this test checks that they are not right by luck of timing.
"""
import itertools

import pytest


def build(K, slack=0, steps=None, drop=()):
    RB = 4 * K + 4 + 2 * slack          # ring rows (clothgpu_api.cu: pl.RB)
    HP = RB // 2
    Kx = (K + 1) // 2                    # set X: stages [0, Kx), set Y: [Kx, K)
    S = steps or (3 * RB)                # enough steps for several trips round the ring
    sets = {0: list(range(0, Kx)), 1: list(range(Kx, K))}

    def pair_group(j):                   # local stage index within the set: {0, 2} -> group pair 0, {1, 3} -> 1
        jl = j if j < Kx else j - Kx
        return (0 if j < Kx else 1, jl & 1)

    ev = {}                              # event -> set of ring slots it touches
    for s in range(S):
        q = s + 1                        # the entry warps deliver pair q (rows 2q, 2q+1) and retire the old occupant
        ev[("E", s)] = {(2 * q) % RB, (2 * q + 1) % RB}
        for j in range(K):
            p = s - 2 * j
            if p >= 1:                   # A_j(0) is the dummy pair: skipped
                ev[("A", j, s)] = {(2 * p) % RB, (2 * p + 1) % RB}
            p = s - 2 * j - 1
            if p >= 0:
                ev[("B", j, s)] = {(2 * p + 1) % RB, (2 * p + 2) % RB}
    names = list(ev)
    idx = {e: i for i, e in enumerate(names)}
    n = len(names)
    succ = [set() for _ in range(n)]

    def edge(a, b):
        if a in idx and b in idx:
            succ[idx[a]].add(idx[b])

    for s in range(S):
        for j, j2 in itertools.product(range(K), repeat=2):
            if pair_group(j) == pair_group(j2):
                edge(("A", j, s), ("B", j2, s))                       # alpha -> beta barrier of a group pair
                edge(("A", j, s), ("A", j2, s + 1))                   # (same threads: program order)
        for st, js in sets.items():
            for j, j2 in itertools.product(js, repeat=2):
                edge(("B", j, s), ("A", j2, s + 1))                   # end-of-step barrier of the set
            if st == 0:
                for j in js:
                    edge(("E", s), ("A", j, s + 1))                   # ... set X's includes the entry warps
                    edge(("B", j, s), ("E", s + 1))
                edge(("E", s), ("E", s + 1))
        if "xdone" not in drop:
            for j, j2 in itertools.product(sets[0], sets[1]):
                edge(("B", j, s), ("A", j2, s + 1))                   # "X has finished step s" -> Y starts s + 1
            for j2 in sets[1]:
                edge(("E", s), ("A", j2, s + 1))                      # (the arrive comes after X's end barrier)
        if "ydone" not in drop:
            for j2 in sets[1]:
                edge(("B", j2, s), ("E", s + 2 + slack))              # "Y has finished step s" -> entry of s + 2
    # transitive closure as bitsets, events in an order compatible with the edges (by step, then E < A < B)
    order = sorted(range(n), key=lambda i: (names[i][-1], "EAB".index(names[i][0])))
    reach = [0] * n
    for i in reversed(order):
        r = 0
        for t in succ[i]:
            r |= (1 << t) | reach[t]
        reach[i] = r
    return dict(K=K, RB=RB, HP=HP, S=S, ev=ev, names=names, idx=idx, reach=reach)


def hb(m, a, b):
    return a in m["idx"] and b in m["idx"] and bool(m["reach"][m["idx"][a]] >> m["idx"][b] & 1)


@pytest.mark.parametrize("slack", [0, 1])
@pytest.mark.parametrize("K", [2, 3, 4, 5, 6, 7, 8])
def test_data_dependencies_are_ordered(K, slack):
    m = build(K, slack)
    S, HP = m["S"], m["HP"]
    for s in range(S):
        for j in range(K):
            if ("A", j, s) in m["idx"]:
                # A_j(p) consumes rows 2p, 2p+1 after sweep j-1: B_{j-1}(p-1) and B_{j-1}(p), or the entry for j = 0
                if j == 0:
                    assert hb(m, ("E", s - 1), ("A", 0, s)), (K, s)
                else:
                    assert hb(m, ("B", j - 1, s - 1), ("A", j, s)), (K, j, s)
                    if ("B", j - 1, s - 2) in m["idx"]:
                        assert hb(m, ("B", j - 1, s - 2), ("A", j, s)), (K, j, s)
            if ("B", j, s) in m["idx"]:
                # B_j(p) consumes A_j(p) (step s - 1) and A_j(p+1) (this step)
                assert hb(m, ("A", j, s), ("B", j, s)), (K, j, s)
                if ("A", j, s - 1) in m["idx"]:
                    assert hb(m, ("A", j, s - 1), ("B", j, s)), (K, j, s)
        # the pair the entry warps write back in step s passed its last colour pass in B_{K-1} of step s - 2 - slack
        qo = s + 1 - HP
        if qo >= 1:
            assert hb(m, ("B", K - 1, s - 2 - slack), ("E", s)), (K, s)
            assert ("B", K - 1, s - 2 - slack) in m["idx"]


@pytest.mark.parametrize("slack", [0, 1])
@pytest.mark.parametrize("K", [2, 3, 4, 5, 6, 7, 8])
def test_unordered_events_never_share_a_ring_slot(K, slack):
    m = build(K, slack)
    names, ev = m["names"], m["ev"]
    by_slot = {}
    for e in names:
        for sl in ev[e]:
            by_slot.setdefault(sl, []).append(e)
    for sl, es in by_slot.items():
        for a, b in itertools.combinations(es, 2):
            if a[0] != "E" and b[0] != "E" and a[1] == b[1] and a[0] == b[0] and a[-1] == b[-1]:
                continue
            assert hb(m, a, b) or hb(m, b, a), f"K={K}: {a} and {b} both touch ring slot {sl} and are not ordered"


def _unordered_conflicts(m):
    by_slot = {}
    for e in m["names"]:
        for sl in m["ev"][e]:
            by_slot.setdefault(sl, []).append(e)
    return sum(1 for es in by_slot.values() for a, b in itertools.combinations(es, 2)
               if not (hb(m, a, b) or hb(m, b, a)))


@pytest.mark.parametrize("drop", ["xdone", "ydone"])
def test_the_model_sees_a_missing_barrier(drop):
    """Sanity of the checker itself: without either of the two producer/consumer barriers between the worker sets,
    unordered events do share ring slots (Y would read rows X has not finished / the entry warps would overwrite rows Y
    still needs)."""
    assert _unordered_conflicts(build(8)) == 0
    assert _unordered_conflicts(build(8, drop=(drop,))) > 0


def test_the_ring_could_be_one_pair_shorter_without_any_drift():
    """4K + 2 ring rows are enough if the entry warps wait for "Y has finished step s - 1" (slack = -1 in the model):
    the kernel spends one more row pair (4K + 4) to let set X run a step ahead of set Y."""
    assert _unordered_conflicts(build(8, slack=-1)) == 0
