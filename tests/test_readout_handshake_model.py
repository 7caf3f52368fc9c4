"""Host-side model of the decode path's hand-overs between the eight unpack warps and the MMA issuer
(bitdelta_b200/csrc/bd_umma.cu: the per-unit loop of the unpack warps, `read_out`, and the MMA issuer's loop):

  * the two warp groups take alternate units; a group's four warps arrive on the unit's A-buffer barrier (one hardware named
    barrier per A buffer, reused every n_abuf units), the MMA issuer syncs on it;
  * a (tile, K run) that is followed by another one is read out one unit LATE: a warp first hands over its first own unit of
    the next run (if it has one), then waits for the run's MMAs (accumulator barrier), reads the accumulators and arrives on
    the "accumulators read" barrier; the MMA issuer syncs on that barrier before the first MMA of the next run.

The model runs each warp's control flow exactly as the kernel's loop does, under random interleavings, and checks: no
deadlock; nobody arrives twice in one phase of a named barrier (which would release the other side early); the MMA issuer
never overwrites accumulators a warp has not read yet; every run is read out exactly once by every warp."""
import random

import pytest


def warp_program(grp: int, n_units: int, kblocks: int, kb0: int):
    """The unpack warp's loop as a list of operations: ('own', unit) | ('readout', run, has_next)."""
    ops = []
    it, rb, kb, run = grp, 0, kb0, 0
    pend = None
    more = n_units > 0
    re = 0
    while True:
        if more:
            re = min(n_units, rb + kblocks - kb)
            if it < re:
                ops.append(("own", it))
                it += 2
        if pend is not None:
            ops.append(("readout",) + pend)
            pend = None
        if not more:
            break
        while it < re:
            ops.append(("own", it))
            it += 2
        pend = (run, re < n_units)
        run += 1
        kb += re - rb
        if kb == kblocks:
            kb = 0
        rb = re
        more = rb < n_units
    return ops


def run_bounds(n_units: int, kblocks: int, kb0: int):
    first, last = {}, {}  # unit -> run it starts / ends
    rb, kb, run = 0, kb0, 0
    while rb < n_units:
        re = min(n_units, rb + kblocks - kb)
        first[rb] = run
        last[re - 1] = run
        kb = (kb + re - rb) % kblocks
        rb, run = re, run + 1
    return first, last, run


def simulate(n_units: int, kblocks: int, kb0: int, n_abuf: int, rng: random.Random) -> None:
    first, last, n_runs = run_bounds(n_units, kblocks, kb0)
    progs = [warp_program(w // 4, n_units, kblocks, kb0) for w in range(8)]
    pc = [0] * 8
    afull = [set() for _ in range(n_abuf)]  # warps arrived in the barrier's current phase
    dempty = set()
    read_by = [set() for _ in range(n_runs)]
    dfull = set()                            # runs whose MMAs are complete
    retired = 0                              # units whose MMAs are complete (in order)
    mma_u, mma_synced_afull, mma_synced_dempty = 0, False, False
    while retired < n_units or any(pc[w] < len(progs[w]) for w in range(8)):
        moves = []
        for w in range(8):
            if pc[w] == len(progs[w]):
                continue
            op = progs[w][pc[w]]
            if op[0] == "own":
                if op[1] < n_abuf or retired > op[1] - n_abuf:  # A buffer free (the stage itself is assumed landed)
                    moves.append(("warp", w))
            elif op[1] in dfull:
                moves.append(("warp", w))
        if mma_u < n_units:
            b = mma_u % n_abuf
            if not mma_synced_afull:
                if len(afull[b]) == 4:
                    moves.append(("mma_afull",))
            elif mma_u in first and first[mma_u] > 0 and not mma_synced_dempty:
                if len(dempty) == 8:
                    moves.append(("mma_dempty",))
            else:
                moves.append(("mma_issue",))
        if retired < mma_u:
            moves.append(("retire",))
        assert moves, f"deadlock at unit {mma_u}: pcs {pc}, retired {retired}"
        m = rng.choice(moves)
        if m[0] == "warp":
            w = m[1]
            op = progs[w][pc[w]]
            pc[w] += 1
            if op[0] == "own":
                u = op[1]
                assert u % 2 == w // 4
                b = u % n_abuf
                assert w not in afull[b], f"warp {w} arrives twice on A-buffer barrier {b}"
                assert mma_u <= u, "unit handed over after the MMA issuer passed it"
                afull[b].add(w)
            else:
                _, run, has_next = op
                assert w not in read_by[run]
                read_by[run].add(w)
                if has_next:
                    assert w not in dempty, f"warp {w} arrives twice on the accumulators-read barrier"
                    dempty.add(w)
        elif m[0] == "mma_afull":
            b = mma_u % n_abuf
            assert all(w // 4 == mma_u % 2 for w in afull[b])
            afull[b].clear()
            mma_synced_afull = True
        elif m[0] == "mma_dempty":
            assert len(read_by[first[mma_u] - 1]) == 8, "accumulators overwritten before every warp has read them"
            dempty.clear()
            mma_synced_dempty = True
        elif m[0] == "mma_issue":
            if mma_u in first and first[mma_u] > 0:
                assert len(read_by[first[mma_u] - 1]) == 8
            mma_u += 1
            mma_synced_afull = mma_synced_dempty = False
        else:
            if retired in last:
                dfull.add(last[retired])
            retired += 1
    assert all(len(r) == 8 for r in read_by)
    assert not dempty and all(not s for s in afull)


@pytest.mark.parametrize("kblocks", [1, 2, 3, 5, 8, 64])
def test_unit_and_readout_handovers(kblocks):
    rng = random.Random(kblocks)
    for n_abuf in (2, 3, 4, 7, 8):
        for n_units in (1, 2, 3, 4, 7, 13, 14, 48, 49):
            for kb0 in sorted({0, kblocks // 2, kblocks - 1}):
                for _ in range(12):
                    simulate(n_units, kblocks, kb0, n_abuf, rng)


def test_warp_program_matches_the_unit_split():
    # every unit is handed over by exactly one group, every run is read out once, and a run's read-out never comes before the
    # group's own units of that run
    for n_units, kblocks, kb0 in [(49, 64, 17), (14, 64, 60), (9, 1, 0), (10, 2, 1), (30, 8, 3)]:
        _, _, n_runs = run_bounds(n_units, kblocks, kb0)
        for grp in (0, 1):
            ops = warp_program(grp, n_units, kblocks, kb0)
            own = [o[1] for o in ops if o[0] == "own"]
            assert own == list(range(grp, n_units, 2))
            runs = [o[1] for o in ops if o[0] == "readout"]
            assert runs == list(range(n_runs))
            assert [o[2] for o in ops if o[0] == "readout"] == [True] * (n_runs - 1) + [False]
