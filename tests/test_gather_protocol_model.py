"""CPU: a randomized-interleaving model of the deferred peer gather (pb_peer_gather push = 3, include/picaso_b200.h).

The device protocol cannot run without GPUs, but its safety argument is arithmetic on step numbers: rotating slots,
the guard a courier waits for before it overwrites a slot on the peers, and the reader contract of the header.  This
model executes the protocol's atomic actions (launch start, the solver's local row store, each courier peer store,
each flag store, reads) in random order under exactly the ordering constraints the device gives:

* launches of one rank run in stream order; inside a launch the courier CTAs (previous step's slab) and the solver
  CTAs (this step's local row) are concurrent;
* the courier of launch s pushes step s - 1 into slot (s - 1) % nbuf of every peer after its own flags from ALL ranks
  reached guard = (s - 1) - (nbuf - 1), and publishes flags[p][r] = s - 1 only after all its stores (fence);
* a rank reads step t after all its flags reached t and - the reader contract - before it launches step t + 2.

The property: every read sees step t in every row (never a stale or a newer slab).  A weaker guard must break it."""
import random

import pytest


def simulate(world, steps, nbuf, seed, guard_slack=0, read_prob=0.5):
    """returns the number of reads that saw a wrong slab; guard_slack > 0 weakens the courier's guard (negative control).
    Every rank reads a random subset of the steps (a rank that reads EVERY step throttles itself through its own
    reads and would hide a missing guard) plus the last one."""
    rng = random.Random(seed)
    reads = [sorted({t for t in range(1, steps) if rng.random() < read_prob} | {steps}) for _ in range(world)]
    next_read = [0] * world                                                # index into reads[r]
    buf = [[[None] * world for _ in range(nbuf)] for _ in range(world)]    # buf[rank][slot][row] = step stored there
    flags = [[0] * world for _ in range(world)]                            # flags[rank][q] = last step q published on rank
    launched = [0] * world                                                 # last launch started on a rank
    # per rank: state of the running launch: None, or dict(step, solver_done, courier=[pending peer stores], published=[...])
    running = [None] * world
    flushed = [False] * world
    bad = 0

    def courier_plan(r, prev):
        return dict(prev=prev, guard=max(0, prev - (nbuf - 1) - guard_slack),
                    stores=[p for p in range(world) if p != r], flags=list(range(world)), waiting=True)

    while True:
        actions = []
        for r in range(world):
            run = running[r]
            if run is None:
                nxt = launched[r] + 1
                # reader contract: the reads of step t (if any) are enqueued before launch t + 2
                pending = reads[r][next_read[r]] if next_read[r] < len(reads[r]) else steps + 9
                if nxt <= steps and pending > nxt - 2:
                    actions.append(("launch", r))
                elif nxt > steps and not flushed[r]:
                    actions.append(("flush", r))
            else:
                if not run["solver_done"]:
                    actions.append(("solver", r))
                c = run["courier"]
                if c is not None:
                    if c["waiting"]:
                        if all(flags[r][q] >= c["guard"] for q in range(world)):
                            actions.append(("courier_go", r))
                    elif c["stores"]:
                        actions.append(("courier_store", r))
                    elif c["flags"]:
                        actions.append(("courier_flag", r))
                if run["solver_done"] and (c is None or (not c["waiting"] and not c["stores"] and not c["flags"])):
                    actions.append(("retire", r))
            if next_read[r] < len(reads[r]) and all(flags[r][q] >= reads[r][next_read[r]] for q in range(world)):
                actions.append(("read", r))
        if not actions:
            break
        kind, r = rng.choice(actions)
        if kind == "launch":
            s = launched[r] = launched[r] + 1
            running[r] = dict(step=s, solver_done=False, courier=courier_plan(r, s - 1) if s > 1 else None)
        elif kind == "flush":          # pb_peer_flush: the last step's slab has no next launch to carry it
            flushed[r] = True
            running[r] = dict(step=steps + 1, solver_done=True, courier=courier_plan(r, steps))
        elif kind == "solver":         # the solver CTAs store this step's row into the LOCAL buffer
            run = running[r]
            if run["step"] <= steps:
                buf[r][run["step"] % nbuf][r] = run["step"]
            run["solver_done"] = True
        elif kind == "courier_go":
            running[r]["courier"]["waiting"] = False
        elif kind == "courier_store":
            c = running[r]["courier"]
            p = c["stores"].pop(rng.randrange(len(c["stores"])))
            buf[p][c["prev"] % nbuf][r] = c["prev"]
        elif kind == "courier_flag":
            c = running[r]["courier"]
            p = c["flags"].pop(rng.randrange(len(c["flags"])))
            flags[p][r] = max(flags[p][r], c["prev"])
        elif kind == "retire":
            running[r] = None
        elif kind == "read":
            t = reads[r][next_read[r]]
            if any(buf[r][t % nbuf][q] != t for q in range(world)):
                bad += 1
            next_read[r] += 1
    assert all(next_read[r] == len(reads[r]) for r in range(world)), "the model deadlocked: %r of %r" % (next_read, reads)
    return bad


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_deferred_gather_never_shows_a_wrong_slab(world):
    for seed in range(150 if world < 8 else 40):
        assert simulate(world, steps=11, nbuf=3, seed=seed) == 0, (world, seed)


def test_more_buffers_are_safe_too():
    for seed in range(60):
        assert simulate(3, steps=13, nbuf=4, seed=seed) == 0


def test_a_weaker_guard_is_caught():
    """negative control: let the courier overwrite a slot one step earlier than wait_step allows - some interleaving
    must then show a reader a slab of the wrong step, or the model is not testing anything"""
    assert sum(simulate(3, steps=11, nbuf=3, seed=seed, guard_slack=1, read_prob=0.3) for seed in range(400)) > 0
