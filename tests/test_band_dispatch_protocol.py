"""Model check of the band kernel's piece dispatch and of the exchange that is fused into it
(csrc/band_kernels.cu: the `while (piece < npieces)` loop, `next_piece`, `sched[0..2]`, `BandEdge`).

Every warp of every group is a little state machine that is advanced in random order, one protocol event at a
time; group barriers are modelled as "all warps of the group have passed the previous phase".  Checked:

* every piece of the list is processed by exactly one group, by all of that group's warps, once;
* the one-piece-ahead fetch: warp 0 writes `next_piece[slot]` at the start of a piece, the group's warps read
  it at the end — never before it was written for THIS piece, never after it was overwritten for the piece after
  next (the two slots alternate);
* the counters are back to zero when the launch ends (the last group re-arms them), also when there are fewer
  pieces than groups, so the next launch — or a CUDA-graph replay — starts from a clean state;
* strips: the epoch flag is posted exactly once per launch, only after EVERY warp of EVERY edge piece has
  fenced its peer stores, and an edge piece never starts before the neighbour's flag of the previous block is
  there; two strips driving each other block after block never deadlock.
Pure Python, no GPU."""
import random

import pytest


class Launch:
    """One kernel launch of one strip: groups x warps walking the dispatch protocol."""

    def __init__(self, rnd, ngroups, nwarps, npieces, edge_pieces, sched, flags_in, flags_out, epoch):
        self.rnd, self.G, self.NW, self.npieces = rnd, ngroups, nwarps, npieces
        self.edge = set(edge_pieces)                 # piece indices that are edge pieces (the last ones of the list)
        self.sched = sched                           # [next, groups_done, edge_warps_done] — lives across launches
        self.flags_in, self.flags_out, self.epoch = flags_in, flags_out, epoch
        self.n_edge_warps = len(self.edge) * nwarps   # what the launcher passes (BandEdge::n_edge_warps)
        self.true_edge_warps = self.n_edge_warps
        self.next_piece = [[None, None] for _ in range(ngroups)]
        self.slot_gen = [[-1, -1] for _ in range(ngroups)]      # which of the group's pieces (ordinal) wrote the slot
        # per warp: (phase, piece, ordinal of the piece within the group, slot parity)
        self.st = {(g, w): ["start", g, 0, 0] for g in range(ngroups) for w in range(nwarps)}
        self.arrived = {}                            # (g, ordinal, barrier) -> warps that arrived
        self.done_by = {}                            # piece -> set of (g, w) that finished it
        self.fenced = 0                              # edge warps that fenced their peer stores
        self.posted = 0
        self.finished_warps = 0

    def runnable(self):
        return [k for k, v in self.st.items() if v[0] != "exit"]

    def barrier(self, g, ordinal, name, w):
        s = self.arrived.setdefault((g, ordinal, name), set())
        s.add(w)
        return len(s) == self.NW

    def passed(self, g, ordinal, name):
        return len(self.arrived.get((g, ordinal, name), ())) == self.NW

    def step(self, key):
        g, w = key
        ph, piece, ordn, pp = self.st[key]
        if ph == "start":
            if piece >= self.npieces:                # while (piece < npieces) fails: leave
                self.st[key][0] = "leave"
                return True
            if w == 0:                               # fetch the piece after this one, write it to the slot
                assert self.slot_gen[g][pp] in (-1, ordn - 2), "slot overwritten before its readers are two pieces behind"
                self.next_piece[g][pp] = self.G + self.sched[0]
                self.sched[0] += 1
                self.slot_gen[g][pp] = ordn
            if piece in self.edge:
                self.st[key][0] = "edge_wait"
            else:
                self.st[key][0] = "work"
            return True
        if ph == "edge_wait":                        # lane 0 polls the neighbour's flag of the previous block
            if self.flags_in is not None and self.flags_in[0] < self.epoch:
                return False                         # blocked (spinning)
            self.st[key][0] = "work"
            return True
        if ph == "work":                             # the piece's bands: at least one full arrive / wait pair
            self.barrier(g, ordn, "a", w)
            self.st[key][0] = "work2"
            return True
        if ph == "work2":
            if not self.passed(g, ordn, "a"):
                return False                         # waiting at the group barrier
            self.done_by.setdefault(piece, set()).add(key)
            if piece in self.edge:                   # fence the peer stores, count, the last one posts
                self.fenced += 1
                self.sched[2] += 1
                if self.sched[2] == self.n_edge_warps:
                    self.sched[2] = 0
                    assert self.fenced == self.true_edge_warps, "flag posted before every edge warp fenced its stores"
                    if self.flags_out is not None:
                        self.flags_out[0] = self.epoch + 1
                    self.posted += 1
            # read the slot: warp 0 wrote it before ITS arrive on barrier "a", which this warp has seen complete
            assert self.slot_gen[g][pp] == ordn, "next_piece read before it was written for this piece"
            self.st[key] = ["start", self.next_piece[g][pp], ordn + 1, pp ^ 1]
            return True
        if ph == "leave":
            had_piece = g < self.npieces
            if had_piece and w == 0:                 # the group that leaves last re-arms the counters
                active = min(self.G, self.npieces)
                self.sched[1] += 1
                if self.sched[1] == active:
                    self.sched[0] = 0
                    self.sched[1] = 0
            self.st[key][0] = "exit"
            self.finished_warps += 1
            return True
        raise AssertionError(ph)

    def finished(self):
        return self.finished_warps == self.G * self.NW

    def check_end(self):
        assert self.sched == [0, 0, 0], f"counters not re-armed: {self.sched}"
        assert sorted(self.done_by) == list(range(self.npieces)), "a piece was skipped"
        for p, who in self.done_by.items():
            groups = {g for g, _ in who}
            assert len(groups) == 1 and len(who) == self.NW, f"piece {p} processed by {who}"
        assert self.posted == (1 if self.edge else 0)


def drive(launches, rnd, max_events=2_000_000):
    """Advance the warps of all running launches in random order; a launch of a strip starts when the previous
    launch of the same strip has finished (stream order).  Returns when every queue is empty."""
    queues = launches                               # list of lists (one list of launch factories per strip)
    running = [None] * len(queues)
    ev = 0
    while True:
        for i, q in enumerate(queues):
            if running[i] is None and q:
                running[i] = q.pop(0)()
        live = [l for l in running if l is not None]
        if not live:
            return
        progressed = False
        order = [(l, k) for l in live for k in l.runnable()]
        rnd.shuffle(order)
        for l, k in order[: max(1, len(order) // 3)]:
            progressed |= l.step(k)
            ev += 1
        if not progressed:                          # everybody chosen was blocked: try all before calling it a deadlock
            progressed = any(l.step(k) for l, k in order)
        assert progressed, "deadlock"
        assert ev < max_events, "livelock"
        for i, l in enumerate(running):
            if l is not None and l.finished():
                l.check_end()
                running[i] = None


@pytest.mark.parametrize("G,NW,npieces", [(4, 3, 37), (6, 2, 6), (8, 2, 3), (5, 4, 1), (3, 3, 100)])
def test_dispatch_covers_every_piece_once_and_rearms(G, NW, npieces):
    for seed in range(6):
        rnd = random.Random(seed)
        sched = [0, 0, 0]
        # three launches in a row on one stream (e.g. a CUDA-graph replay): the counters must come back clean each time
        drive([[lambda: Launch(rnd, G, NW, npieces, [], sched, None, None, 0)] * 3], rnd)
        assert sched == [0, 0, 0]


@pytest.mark.parametrize("G,NW,n_int,n_edge", [(4, 3, 20, 4), (6, 2, 2, 6), (3, 2, 0, 5), (8, 2, 30, 2)])
def test_fused_exchange_between_two_strips(G, NW, n_int, n_edge):
    """Two strips, five blocks each.  Block b of a strip waits for the neighbour's epoch >= b (posted by the
    neighbour's block b-1) and posts epoch b+1 — only after all of its edge warps have fenced."""
    for seed in range(6):
        rnd = random.Random(100 + seed)
        npieces = n_int + n_edge
        edge = list(range(n_int, npieces))           # the edge bands are the LAST pieces of the list
        flag_a, flag_b = [0], [0]                    # flag_x: written by the neighbour, polled by strip x
        sa, sb = [0, 0, 0], [0, 0, 0]
        blocks = 5
        qa = [(lambda b=b: Launch(rnd, G, NW, npieces, edge, sa, flag_a, flag_b, b)) for b in range(blocks)]
        qb = [(lambda b=b: Launch(rnd, G, NW, npieces, edge, sb, flag_b, flag_a, b)) for b in range(blocks)]
        drive([qa, qb], rnd)
        assert flag_a == [blocks] and flag_b == [blocks]


def test_model_catches_a_flag_posted_by_the_first_edge_warp():
    """Sanity of the model: posting as soon as ONE edge warp is done must trip the fence check."""
    class Broken(Launch):
        def __init__(self, *a):
            super().__init__(*a)
            self.n_edge_warps = 1
    rnd = random.Random(7)
    with pytest.raises(AssertionError, match="fenced"):
        for _ in range(20):
            drive([[lambda: Broken(rnd, 4, 3, 12, [8, 9, 10, 11], [0, 0, 0], None, [0], 0)]], rnd)
