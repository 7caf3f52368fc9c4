"""Host-side model of the fused kernel's per-stage barrier protocol (bitdelta_b200/csrc/bd_umma.cu, "one-time setup" and the
MMA issuer's commits): the barrier of stage s = u % stages is armed with an arrival count of TWO and collects, for unit u,

  * the producer's arrive.expect_tx for unit u (issued once the stage's previous user, unit u - stages, has retired), and
  * the MMA commit of unit u - n_abuf (the previous user of unit u's A buffer), or the set-up pre-arrival for u < n_abuf,

so one wait tells the sync warp both "stage landed" and "A buffer free".  The kernel relies on n_abuf <= stages for the commit
never to fall into the stage's PREVIOUS phase; the host plan clamps n_abuf accordingly.  The model replays random legal event
orders and checks that every arrival lands in the phase of the unit it is meant for -- and that the property really does
break for n_abuf > stages, i.e. that the clamp is needed."""
import random

import pytest


class PhaseError(AssertionError):
    pass


def simulate(stages: int, n_abuf: int, n_units: int, rng: random.Random) -> None:
    # barrier s: [phase index, arrivals seen in that phase]; phase p of barrier s belongs to unit u = p * stages + s
    bars = [[0, 0] for _ in range(stages)]

    def arrive(s: int, for_unit: int, who: str) -> None:
        phase, _ = bars[s]
        if phase * stages + s != for_unit:
            raise PhaseError(f"{who}: arrival for unit {for_unit} landed in the phase of unit {phase * stages + s}")
        bars[s][1] += 1
        if bars[s][1] == 2:  # phase complete: the sync warp may release the unit; the barrier moves on
            completed.add(for_unit)
            bars[s] = [phase + 1, 0]

    completed: set[int] = set()
    for u in range(min(n_abuf, n_units, stages)):  # set-up: the first n_abuf units find their A buffer free
        arrive(u, u, "pre-arrival")
    produced = released = issued = retired = 0  # units the producer / sync warp / MMA issuer / tensor pipe are done with
    while retired < n_units:
        moves = []
        if produced < n_units and (produced < stages or retired > produced - stages):
            moves.append("produce")  # stage free: its previous user's MMAs have retired (bar_empty)
        if released < n_units and released in completed:
            moves.append("release")  # the sync warp's ONE wait per unit
        if issued < released:
            moves.append("issue")    # unpack + MMA issue, in unit order
        if retired < issued:
            moves.append("retire")   # tcgen05.commit arrivals fire when the unit's MMAs complete, in order
        assert moves, f"deadlock: produced {produced} released {released} issued {issued} retired {retired}"
        m = rng.choice(moves)
        if m == "produce":
            arrive(produced % stages, produced, "producer")
            produced += 1
        elif m == "release":
            released += 1
        elif m == "issue":
            issued += 1
        else:
            nxt = retired + n_abuf  # next user of this unit's A buffer
            if nxt < n_units:
                arrive(nxt % stages, nxt, "MMA commit")
            retired += 1
    assert released == n_units


@pytest.mark.parametrize("stages", range(3, 9))
def test_every_arrival_lands_in_its_units_phase(stages):
    rng = random.Random(stages)
    for n_abuf in range(2, stages + 1):
        for n_units in (1, 2, n_abuf, stages, stages + 1, 3 * stages + 1, 41):
            for _ in range(40):
                simulate(stages, n_abuf, n_units, rng)


def test_more_a_buffers_than_stages_breaks_the_protocol():
    # the reason for the host-side clamp n_abuf <= stages: unit u - n_abuf's commit would then target a stage whose current
    # phase still belongs to an EARLIER unit than u
    rng = random.Random(0)
    broken = 0
    for _ in range(200):
        try:
            simulate(4, 6, 30, rng)
        except (PhaseError, AssertionError):
            broken += 1
    assert broken == 200
