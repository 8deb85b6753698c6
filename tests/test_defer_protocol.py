"""The deferral protocol of direct_kernel<.., DEFER> (csrc/rt_kernels.cuh), modelled in Python and run under a random scheduler.

A unit of the deferring variant is a batch of 32 samples with ALL its lights.  A light whose shadow beam is "heavy" is
not processed where it is met: the unit writes its sums so far as the record after light g-1, reserves list slots for
light g AND every later light of the batch with one compare-and-swap (all or nothing; no room: the light is processed
in place), lists light g as M sub-tasks, and goes on through its remaining lights as the per-light units of its batch
would -- an ordinary later light parks its M contributions in its slot and reports them complete, a heavy one is
listed too.  Warps that find the unit queue dry work the sub-tasks off (tickets in publication order); whoever
completes the last missing piece of a slot adds its contributions to the predecessor's record (finish_chain: looks at
the record AFTER done[] is complete, and the finisher of the predecessor looks at done[idx + 1] AFTER it has written
its record, so at least one of the two sees the other; `claim` makes it exactly one) and carries on into the next slot.

The model keeps exactly the kernel's shared words (n_heavy, n_ready, ready[], next_task, done[], claim[], the records)
and yields to a random scheduler between any two accesses to them.  Checked for random interleavings:
  * no warp ever waits for anything but the publication of a ticket's list entry (a bounded, always-satisfied wait);
  * every batch's lights end up added exactly once, in light order, whatever the capacity of the list;
  * the launch terminates.
Host logic only -- the kernel itself is covered by tests/test_gpu_parity.py (test_deferred_units_are_bit_identical,
test_deferral_never_blocks_when_the_list_fills_up, test_more_lights_than_a_unit_takes)."""
import random

import pytest


class Launch:
    def __init__(self, n_batches, L, M, cap, heavy_p, seed):
        rng = random.Random(seed)
        self.nb, self.L, self.M, self.cap = n_batches, L, M, cap
        self.heavy = [[rng.random() < heavy_p for _ in range(L)] for _ in range(n_batches)]
        self.next_unit = 0
        self.n_heavy = 0
        self.n_ready = 0
        self.ready = {}
        self.unit_of = {}
        self.next_task = 0
        self.done = {}
        self.claim = {}
        self.slab = {}
        self.rec = {}      # (light, batch) -> ("final", [lights added so far]) | ("deferred", slot)
        self.sample = {}   # batch -> [lights in the order they were added]
        self.warps_done = 0
        self.in_place = 0
        self.listed = 0


def finish_chain(s, idx, g, b):
    base = []
    if g > 0:
        r = s.rec.get((g - 1, b))
        yield
        if r is None or r[0] != "final":
            return  # the predecessor's finisher will come here
        base = r[1]
    if s.claim.get(idx):
        return
    s.claim[idx] = 1  # (atomicCAS: test and set in one step)
    yield
    while True:
        assert s.done.get(idx, 0) == s.M and len(s.slab[idx]) == s.M, "a slot is finished before its contributions are complete"
        sums = base + [g]
        if g + 1 != s.L:
            s.rec[(g, b)] = ("final", sums)
        else:
            assert b not in s.sample, "a sample is written twice"
            s.sample[b] = sums
        yield
        if g + 1 == s.L:
            return
        d = s.done.get(idx + 1, 0)
        yield
        if d != s.M or s.claim.get(idx + 1):
            return
        s.claim[idx + 1] = 1
        yield
        idx, g, base = idx + 1, g + 1, sums


def warp(s, n_warps):
    # ---- the unit queue
    while True:
        b = s.next_unit
        s.next_unit += 1  # (atomicAdd)
        yield
        if b >= s.nb:
            break
        sums, parked, slot = [], False, None
        for g in range(s.L):
            heavy = s.heavy[b][g]
            if heavy or parked:
                idx = slot
                if not parked:
                    idx = None
                    old = s.n_heavy
                    yield
                    while old + (s.L - g) <= s.cap:
                        if s.n_heavy == old:  # (atomicCAS)
                            s.n_heavy = old + (s.L - g)
                            idx = old
                            break
                        old = s.n_heavy
                        yield
                    yield
                    if idx is not None and g > 0:
                        s.rec[(g - 1, b)] = ("final", list(sums))  # the sums so far: what the list continues from
                        yield
                if idx is not None and heavy:  # defer_unit
                    if g + 1 != s.L:
                        s.rec[(g, b)] = ("deferred", idx)
                        yield
                    s.unit_of[idx] = (g, b)
                    yield
                    r = s.n_ready
                    s.n_ready += 1  # (atomicAdd)
                    yield
                    s.ready[r] = idx
                    yield
                    s.listed += 1
                    parked, slot = True, idx + 1
                    continue
                if idx is not None:  # an ordinary light behind a listed one: contributions now, sums later
                    if g + 1 != s.L:
                        s.rec[(g, b)] = ("deferred", idx)
                        yield
                    s.slab[idx] = list(range(s.M))
                    yield
                    s.done[idx] = s.done.get(idx, 0) + s.M  # (atomicAdd)
                    yield
                    yield from finish_chain(s, idx, g, b)
                    parked, slot = True, idx + 1
                    continue
                s.in_place += 1
            sums.append(g)  # processed in place
            yield
        if not parked:
            assert b not in s.sample
            s.sample[b] = sums
            yield
    # ---- the queue is dry: sub-tasks of the listed lights
    s.warps_done += 1
    yield
    while True:
        ticket = s.next_task
        s.next_task += 1  # (atomicAdd)
        yield
        pos = ticket // s.M
        while True:
            final_list = s.warps_done >= n_warps
            yield
            if pos < s.n_ready:
                break
            if final_list:
                return  # the launch has no such sub-task
            yield "waiting for a list entry"
        while pos not in s.ready:
            yield "waiting for a list entry"  # (counted, being written)
        idx = s.ready[pos]
        g, b = s.unit_of[idx]
        s.slab.setdefault(idx, []).append(ticket - pos * s.M)
        yield
        s.done[idx] = s.done.get(idx, 0) + 1
        finished = s.done[idx]
        yield
        if finished == s.M:
            yield from finish_chain(s, idx, g, b)


def run(n_batches, L, M, cap, heavy_p, n_warps, seed):
    s = Launch(n_batches, L, M, cap, heavy_p, seed)
    rng = random.Random(seed + 1)
    warps = [warp(s, n_warps) for _ in range(n_warps)]
    live = list(range(n_warps))
    idle_rounds = 0
    while live:
        i = rng.choice(live)
        try:
            what = next(warps[i])
        except StopIteration:
            live.remove(i)
            continue
        # a warp may only ever wait for a list entry; if every live warp does, nobody can publish one: a deadlock
        idle_rounds = idle_rounds + 1 if what == "waiting for a list entry" else 0
        assert idle_rounds < 20000, "every warp is waiting"
    assert sorted(s.sample) == list(range(n_batches)), "a sample was never written"
    assert all(v == list(range(L)) for v in s.sample.values()), "lights added out of order or twice"
    return s


@pytest.mark.parametrize("cap", [10**6, 60, 7, 0])
def test_deferral_with_units_of_all_lights_adds_every_light_once_and_in_order(cap):
    listed = in_place = 0
    for seed in range(12):
        s = run(n_batches=90, L=5, M=3, cap=cap, heavy_p=0.25, n_warps=rng_warps(seed), seed=seed)
        listed += s.listed
        in_place += s.in_place
    if cap >= 10**6:
        assert listed > 0 and in_place == 0
    if cap == 0:
        assert listed == 0 and in_place > 0
    if cap in (60, 7):
        assert listed > 0 and in_place > 0  # both the list and the list-full path are exercised


def rng_warps(seed):
    return [3, 8, 17, 40][seed % 4]


def test_every_light_heavy_and_a_single_light():
    run(n_batches=40, L=5, M=4, cap=10**6, heavy_p=1.0, n_warps=9, seed=3)
    run(n_batches=40, L=1, M=4, cap=10**6, heavy_p=0.5, n_warps=9, seed=4)
    run(n_batches=40, L=8, M=2, cap=50, heavy_p=0.5, n_warps=30, seed=5)


def test_the_model_sees_a_protocol_without_the_cascade(monkeypatch):
    """Sanity of the model itself: a finish_chain that does its own slot but does not carry on into the next one leaves
    samples unwritten whenever a later light was complete before its predecessor -- and the model says so."""
    import sys

    this = sys.modules[__name__]

    def no_cascade(s, idx, g, b):
        base = []
        if g > 0:
            r = s.rec.get((g - 1, b))
            yield
            if r is None or r[0] != "final":
                return
            base = r[1]
        if s.claim.get(idx):
            return
        s.claim[idx] = 1
        yield
        sums = base + [g]
        if g + 1 != s.L:
            s.rec[(g, b)] = ("final", sums)
        else:
            s.sample[b] = sums
        yield

    monkeypatch.setattr(this, "finish_chain", no_cascade)
    failures = 0
    for seed in range(6):
        try:
            run(60, 5, 3, 10**6, 0.3, 8, seed)
        except AssertionError:
            failures += 1
    assert failures > 0
