"""Model check of the row-streaming kernel's neighbour synchronisation (csrc/stream_kernels.cu,
TWS_STREAM_WAIT 4): one mbarrier per (row slot, half-pass, ring-turn parity), waited on by PHASE
PARITY only.  A parity wait is exact only if the awaited phase is the one in progress or the one just
completed; this test replays the protocol under random and starved warp schedules and checks that a
wait never passes before the awaited half-pass has really happened (no false positive) and that the
pipeline always drains (no deadlock).  Pure Python, no GPU."""
import random

import pytest


def simulate(NW, K, rows_out, seed, starve):
    rnd = random.Random(seed)
    NHP = 2 * K + 1
    nrows = rows_out + 4 * K                      # 2K warm-up rows above, 2K feeder rows below
    yb = 2 * K + rows_out
    phase = [[[0, 0] for _ in range(NHP)] for _ in range(NW)]   # completed phases of evt[slot][e][turn & 1]
    done = set()
    state = [{"idx": w, "s": 0} for w in range(NW)]
    slow = set(rnd.sample(range(NW), k=max(1, NW // 4))) if starve else set()

    def smax(idx):                                # rows below the piece stop early (they only feed the rows above)
        return 2 * K if idx < yb else 2 * K - (idx - yb + 1)

    def try_wait(slot, e, turn):                  # mbarrier.try_wait.parity: true iff the current phase's parity differs
        return (phase[slot][e][turn & 1] & 1) != ((turn >> 1) & 1)

    while True:
        live = [w for w in range(NW) if state[w]["idx"] < nrows]
        if not live:
            return "ok"
        rnd.shuffle(live)
        if starve and rnd.random() < 0.9:
            fast = [w for w in live if w not in slow]
            live = fast + [w for w in live if w in slow]     # starved warps only run when nobody else can
        progressed = False
        for w in live:
            st = state[w]
            idx, s = st["idx"], st["s"]
            if s > smax(idx):
                st["idx"] += NW
                st["s"] = 0
                progressed = True
                break
            if s >= 1:
                wup, wdn = (w - 1) % NW, (w + 1) % NW
                ok_u = idx == 0 or try_wait(wup, s - 1, (idx - 1) // NW)
                ok_d = try_wait(wdn, s - 1, (idx + 1) // NW)
                if not (ok_u and ok_d):
                    continue
                if idx > 0 and (idx - 1, s - 1) not in done:
                    return f"false positive on the row above: row {idx} half-pass {s}"
                if (idx + 1, s - 1) not in done:
                    return f"false positive on the row below: row {idx} half-pass {s}"
            done.add((idx, s))
            phase[w][s][(idx // NW) & 1] += 1     # arrive
            st["s"] += 1
            progressed = True
            if rnd.random() < 0.7:
                break
        if not progressed:
            return "deadlock"


@pytest.mark.parametrize("NW,K", [(32, 1), (32, 2), (32, 3), (32, 4), (16, 2), (16, 4), (9, 4), (5, 2), (3, 1)])
@pytest.mark.parametrize("starve", [False, True], ids=["fair", "starved"])
def test_parity_waits_are_exact_and_the_pipeline_drains(NW, K, starve):
    for rows_out in (1, 2, NW - 1, NW, 3 * NW + 5, 200):
        for seed in range(6):
            assert simulate(NW, K, rows_out, seed, starve) == "ok", (NW, K, rows_out, seed)


def test_single_barrier_set_is_not_enough():
    """With ONE barrier per (slot, half-pass) the last slot's first-turn wait on slot 0 is ambiguous
    (slot 0 may be a whole turn behind or ahead): the model must find that, or it proves nothing."""
    def simulate_one_set(NW, K, rows_out, seed):
        rnd = random.Random(seed)
        NHP = 2 * K + 1
        nrows = rows_out + 4 * K
        yb = 2 * K + rows_out
        phase = [[0] * NHP for _ in range(NW)]
        done = set()
        state = [{"idx": w, "s": 0} for w in range(NW)]
        smax = lambda idx: 2 * K if idx < yb else 2 * K - (idx - yb + 1)
        while True:
            live = [w for w in range(NW) if state[w]["idx"] < nrows]
            if not live:
                return "ok"
            rnd.shuffle(live)
            progressed = False
            for w in live:
                st = state[w]
                idx, s = st["idx"], st["s"]
                if s > smax(idx):
                    st["idx"] += NW; st["s"] = 0; progressed = True
                    break
                if s >= 1:
                    wup, wdn = (w - 1) % NW, (w + 1) % NW
                    ok_u = idx == 0 or (phase[wup][s - 1] & 1) != (((idx - 1) // NW) & 1)
                    ok_d = (phase[wdn][s - 1] & 1) != (((idx + 1) // NW) & 1)
                    if not (ok_u and ok_d):
                        continue
                    if (idx > 0 and (idx - 1, s - 1) not in done) or (idx + 1, s - 1) not in done:
                        return "false positive"
                done.add((idx, s)); phase[w][s] += 1; st["s"] += 1; progressed = True
                if rnd.random() < 0.7:
                    break
            if not progressed:
                return "deadlock"
    results = {simulate_one_set(8, 2, 100, seed) for seed in range(40)}
    assert results - {"ok"}, "the single-set protocol should fail under some schedule"
