"""Randomised interleaving check of the V-tile ring of k_blk_flush6 (ellp_b200/csrc/blocked.cuh): no producer warp -- the ring slot of
step s is refilled with the tile of step s + stages by whichever of the 16 consumer warps is the LAST to finish the DMMAs of step s
(shared-memory arrival counter per slot, reset by that warp before it re-arms the slot's `full` mbarrier).  compute-sanitizer is closed
on this pool, so the protocol is restated here as a state machine and run under random schedules; the invariants are the two hazards
of a ring: a tile is never overwritten while a warp may still read it, and a warp never reads a slot before the tile of ITS step landed.
Pure Python, no GPU."""
import random

import pytest


def run(nwarps, stages, nsteps, rng, copy_delay_max):
    full_phase = [0] * stages         # completed phases of full[slot] (mbarrier phase counter)
    pending_copy = [None] * stages    # (tile, remaining delay) of an armed, not yet landed bulk copy
    slot_tile = [None] * stages       # tile currently in the slot's shared memory
    arrived = [0] * stages
    readers = [set() for _ in range(stages)]  # warps currently reading the slot (between their wait and their arrival)
    pc = [0] * nwarps                 # step each warp works on
    state = ["wait"] * nwarps         # wait -> read -> (arrive, maybe refill) -> wait of the next step
    for t in range(min(stages, nsteps)):  # initial fill by warp 0 (before any consumer waits: it is in program order of warp 0 only,
        pending_copy[t] = (t, rng.randint(0, copy_delay_max))  # the others simply find the barrier incomplete)
    done = 0
    guard = 0
    while done < nwarps:
        guard += 1
        assert guard < 10_000_000, "deadlock"
        # copies make progress independently of the warps
        for sl in range(stages):
            if pending_copy[sl] is not None:
                tile, d = pending_copy[sl]
                if d == 0:
                    assert not readers[sl], f"tile {tile} lands in slot {sl} while warps {readers[sl]} still read it"
                    slot_tile[sl] = tile
                    full_phase[sl] += 1
                    pending_copy[sl] = None
                else:
                    pending_copy[sl] = (tile, d - 1)
        w = rng.randrange(nwarps)
        if state[w] == "done":
            continue
        s = pc[w]
        sl = s % stages
        if state[w] == "wait":
            # mbar_wait(&full[slot], (s / stages) & 1): try_wait.parity succeeds when the barrier's CURRENT phase has the other parity
            # (i.e. the phase with the given parity completed); the assert below also covers aliasing of phases two apart
            if (full_phase[sl] & 1) != ((s // stages) & 1):
                assert slot_tile[sl] == s, f"warp {w} at step {s} would read tile {slot_tile[sl]}"
                readers[sl].add(w)
                state[w] = "read"
        elif state[w] == "read":
            assert slot_tile[sl] == s
            readers[sl].discard(w)   # the DMMAs consumed the fragments: this warp no longer reads the slot
            arrived[sl] += 1
            if arrived[sl] == nwarps:          # atomicAdd returned nwarps - 1: this warp is the last one
                nt = s + stages
                if nt < nsteps:
                    arrived[sl] = 0            # reset BEFORE re-arming: nobody arrives again until the copy below completed a phase
                    assert pending_copy[sl] is None
                    pending_copy[sl] = (nt, rng.randint(0, copy_delay_max))
            pc[w] += 1
            if pc[w] == nsteps:
                state[w] = "done"
                done += 1
            else:
                state[w] = "wait"
    return True


@pytest.mark.parametrize("seed", range(30))
def test_last_arriver_refill_never_overwrites_a_live_tile_and_never_serves_a_stale_one(seed):
    rng = random.Random(seed)
    for _ in range(20):
        nwarps = rng.choice([1, 2, 4, 16])
        stages = rng.choice([1, 2, 3])
        nsteps = rng.randint(1, 40)
        assert run(nwarps, stages, nsteps, rng, copy_delay_max=rng.choice([0, 3, 50, 400]))
