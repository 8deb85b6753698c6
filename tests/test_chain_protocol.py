"""The windowed chain of direct_kernel (csrc/rt_kernels.cuh), modelled in Python and run under a random scheduler.

direct_kernel's work units are (window w, light group g, batch column b), handed out in that order to persistent
warps.  A lane that hit something continues the running light sums its predecessor unit (w, g-1, b) left in a tagged
record; records live in two generations per column (slot = g & 1) and are reused by the next window.  The slot reuse
is made safe by one word per column: the LAST unit of a column publishes its tag in done[b]; the FIRST unit of the
next window -- and, since round 2, also its LAST unit -- waits until done[b] has reached the previous window's tag.

The model keeps exactly the kernel's waits and stores (lanes that missed take no part in the record chain; a unit none
of whose lanes hit has no record dependency at all) and checks, for random interleavings of R warps:
  * progress: no state in which every warp waits (the round-2 GPU runs hung with the first version of the protocol:
    a last unit without hits published at once and overwrote the tag the first unit of its window still waited for);
  * safety: no record is overwritten before the lane it was written for has read it.
Host logic only -- the kernel itself is covered by tests/test_gpu_parity.py::test_chained_records_live_in_a_bounded_window."""
import random

import pytest


def run(n_batches, W, G, R, seed, last_unit_waits, monotonic, lanes=2, frame_tag=1):
    rng = random.Random(seed)
    window_units = W * G
    n_units = n_batches * G
    hit = {(bt, ln): rng.random() < 0.55 for bt in range(n_batches) for ln in range(lanes)}
    for bt in range(0, n_batches, 7):  # whole columns without a hit (the black border of the frame)
        for ln in range(lanes):
            hit[(bt, ln)] = False
    done, rec, unread = {}, {}, {}
    nxt, finished = 0, 0
    warps = [None] * R

    def decode(unit):
        win, in_win = divmod(unit, window_units)
        wb = min(W, n_batches - win * W)
        g, b = divmod(in_win, wb)
        return win, g, b

    while finished < n_units:
        progressed = False
        order = list(range(R))
        rng.shuffle(order)
        for i in order:
            w = warps[i]
            if w is None:
                if nxt < n_units:
                    win, g, b = decode(nxt)
                    warps[i] = dict(win=win, g=g, b=b, stage=0, delay=rng.choice([0, 0, 1, 4, 50]), batch=win * W + b)
                    nxt += 1
                    progressed = True
                continue
            tag_base = frame_tag + w["win"] * G
            last = w["g"] + 1 == G
            if w["stage"] == 0:  # the done[] wait
                if w["win"] > 0 and G > 1 and (w["g"] == 0 or (last and last_unit_waits)):
                    v, want = done.get(w["b"], 0), tag_base - 1
                    if not ((v - want >= 0) if monotonic else (v == want)):
                        continue
                w["stage"] = 1
                progressed = True
            elif w["stage"] == 1:  # every lane that hit waits for its predecessor's record
                if w["g"] > 0:
                    want = tag_base + w["g"] - 1
                    slot = (w["g"] & 1) ^ 1
                    if any(hit[(w["batch"], ln)] and rec.get((slot, w["b"], ln)) != want for ln in range(lanes)):
                        continue
                    for ln in range(lanes):
                        if hit[(w["batch"], ln)]:
                            unread.pop((slot, w["b"], ln), None)  # read
                w["stage"] = 2
                progressed = True
            else:
                if w["delay"] > 0:
                    w["delay"] -= 1
                    progressed = True
                    continue
                if not last:
                    for ln in range(lanes):
                        if hit[(w["batch"], ln)]:
                            key = (w["g"] & 1, w["b"], ln)
                            if key in unread:
                                return "OVERWRITE"
                            rec[key] = tag_base + w["g"]
                            unread[key] = True
                if last and G > 1 and W < n_batches:
                    done[w["b"]] = tag_base + w["g"]
                warps[i] = None
                finished += 1
                progressed = True
        if not progressed:
            return "DEADLOCK"
    return "ok"


@pytest.mark.parametrize("W,R", [(20, 45), (24, 12), (16, 64), (50, 20)])
def test_windowed_chain_makes_progress_and_never_overwrites_an_unread_record(W, R):
    for seed in range(10):
        assert run(310, W, 5, R, seed, last_unit_waits=True, monotonic=True) == "ok"
    assert run(310, W, 2, R, 3, True, True) == "ok" and run(310, W, 3, R, 4, True, True) == "ok"


def test_the_first_version_of_the_protocol_hangs():
    """What the GPU showed (watchdog code 6, RT_CHAIN_WINDOW probe in profiles/README.md): without the last unit's wait
    a column without hits publishes its tag at once, and the first unit of its own window waits forever."""
    results = [run(310, 20, 5, 45, seed, last_unit_waits=False, monotonic=False) for seed in range(10)]
    assert "DEADLOCK" in results
