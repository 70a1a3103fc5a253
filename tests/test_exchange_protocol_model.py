"""Model check (CPU, pure Python) of the tile-to-tile halo exchange protocol of gb-25_b200/csrc/gb25_exchange.cu
(launch_fill_halo_dist, x phase): every fill PUSHES the packed edge columns into the neighbours' column inbox — double
buffered by the parity of a per-lane column-phase counter —, publishes a per-lane sequence number in the neighbours' flag
words (one word per sender slot), waits for the numbers of the tiles it receives from (>=), and unpacks its own inbox.  There
are no acknowledgements.  Two lanes run concurrently on two streams of every tile (lane 1: T, S under the barotropic solve,
forked from and joined to the main stream every step) with their own counters, flag words and inbox slots.

Every (tile, lane) is a FIFO of operations; a randomised, unfair scheduler executes any operation whose wait is satisfied.  A
strip that is overwritten before it was unpacked, a flag word shared between senders or lanes, or a missing ordering shows up
as an unpacked strip that carries the wrong (sender, lane, fill) tag, or as a deadlock.  The negative controls break the
protocol on purpose (single-buffered inbox; lanes sharing the inbox slots; lanes sharing the flag words) and must be caught —
so the model can see the hazards it is there to exclude.  Host-side logic: no GPU."""
import random

import pytest

W, E = 0, 1          # my inbox halves: strips that came from the west / from the east neighbour; also the flag slots


class Violation(Exception):
    pass


def run(Rx, nsteps, seed, double_buffered=True, lane_slots=True, lane_flags=True, row_phase_prob=0.5):
    rng = random.Random(seed)
    tiles = range(Rx)
    nb = {t: ((t - 1) % Rx, (t + 1) % Rx) for t in tiles}                 # (west, east) neighbour
    # inbox[tile][parity][half][slot] = tag; slot 0: lane 0's fields, slot 1: lane 1's
    inbox = {t: [[[None, None], [None, None]] for _ in range(2)] for t in tiles}
    flags = {t: [[0, 0], [0, 0]] for t in tiles}                            # flags[tile][lane][sender slot]
    done_fork = {t: -1 for t in tiles}                                      # last step whose fork event main has recorded
    done_join = {t: -1 for t in tiles}                                      # last step whose lane-1 work has finished
    # the program is the same on every tile (SPMD): which fills of lane 0 have a row phase is decided once
    plan = [[rng.random() < row_phase_prob for _ in range(2)] for _ in range(nsteps)]

    def fill_ops(t, lane, step, k, with_rows, counters):
        """One launch_fill_halo_dist: optional row phase (advances seq only), then the column phase."""
        ops = []
        if with_rows:
            counters["seq"] += 1
            s = counters["seq"]
            # (rows go to other tiles in the real code; here only the numbering matters: the same flag words carry it)
            ops.append(("signal", lane, s))
            ops.append(("wait", lane, s))
        counters["seq"] += 1
        counters["xseq"] += 1
        s, par = counters["seq"], (counters["xseq"] & 1) if double_buffered else 0
        tag = (lane, step, k)
        ops.append(("push", lane, par, tag))
        ops.append(("signal", lane, s))
        ops.append(("wait", lane, s))
        ops.append(("unpack", lane, par, tag))
        return ops

    prog = {(t, l): [] for t in tiles for l in (0, 1)}
    for t in tiles:
        c0, c1 = {"seq": 0, "xseq": 0}, {"seq": 0, "xseq": 0}
        for step in range(nsteps):
            prog[t, 0].append(("fork", step))
            prog[t, 1].append(("await_fork", step))
            prog[t, 1] += fill_ops(t, 1, step, 0, False, c1)            # T, S fill on the second stream
            prog[t, 1].append(("joined", step))
            for k in range(2):                                          # barotropic-side fill(s) and the u, v fill on main
                prog[t, 0] += fill_ops(t, 0, step, k, plan[step][k], c0)
            prog[t, 0].append(("await_join", step))
    pc = {key: 0 for key in prog}

    def ready(t, l):
        if pc[t, l] >= len(prog[t, l]):
            return False
        op = prog[t, l][pc[t, l]]
        if op[0] == "wait":
            fl = flags[t][op[1] if lane_flags else 0]
            return fl[W] >= op[2] and fl[E] >= op[2]
        if op[0] == "await_fork":
            return done_fork[t] >= op[1]
        if op[0] == "await_join":
            return done_join[t] >= op[1]
        return True

    def execute(t, l):
        op = prog[t, l][pc[t, l]]
        pc[t, l] += 1
        west, east = nb[t]
        if op[0] == "push":
            _, lane, par, tag = op
            slot = lane if lane_slots else 0
            inbox[east][par][W][slot] = (t,) + tag                       # my last columns -> east tile's "from the west"
            inbox[west][par][E][slot] = (t,) + tag
        elif op[0] == "signal":
            _, lane, s = op
            fl = lane if lane_flags else 0
            flags[east][fl][W] = max(flags[east][fl][W], s) if not lane_flags else s
            flags[west][fl][E] = max(flags[west][fl][E], s) if not lane_flags else s
        elif op[0] == "unpack":
            _, lane, par, tag = op
            slot = lane if lane_slots else 0
            if inbox[t][par][W][slot] != (west,) + tag or inbox[t][par][E][slot] != (east,) + tag:
                raise Violation(f"tile {t} lane {lane} fill {tag}: unpacked {inbox[t][par][W][slot]} / {inbox[t][par][E][slot]}")
        elif op[0] == "fork":
            done_fork[t] = op[1]
        elif op[0] == "joined":
            done_join[t] = op[1]

    favourite = None
    while any(pc[key] < len(prog[key]) for key in prog):
        cand = [key for key in prog if ready(*key)]
        if not cand:
            raise Violation("deadlock")
        # unfair: keep running one (tile, lane) for a random burst, so that it gets as far ahead as the protocol allows
        if favourite not in cand or rng.random() < 0.2:
            favourite = rng.choice(cand)
        execute(*favourite)
    return True


@pytest.mark.parametrize("Rx", [2, 4, 8])
def test_exchange_protocol_holds_under_random_unfair_scheduling(Rx):
    for seed in range(60):
        assert run(Rx, nsteps=6, seed=seed)


@pytest.mark.parametrize("broken", ["single_buffered", "shared_slots", "shared_flags"])
def test_model_catches_the_hazards_the_protocol_is_built_to_exclude(broken):
    kw = {"single_buffered": dict(double_buffered=False), "shared_slots": dict(lane_slots=False),
          "shared_flags": dict(lane_flags=False)}[broken]
    caught = 0
    for Rx in (2, 4):
        for seed in range(60):
            try:
                run(Rx, nsteps=6, seed=seed, **kw)
            except Violation:
                caught += 1
    assert caught > 0, broken
