"""Model check of the resident kernel's rim exchange (csrc/resident_kernels.cu): every block publishes the rim of
step t as words {value, tag = t} into mailbox parity t & 1 and, before the late outflow pass of step t + 1, polls its
neighbours' words until their tag equals t.  There is no flag and no barrier between blocks, and only TWO parities.

Claims checked here, under random and adversarial (starved) block schedules, for 1-D rings, 2-D grids with the kernel's
8-neighbourhood, and frames that continue the step count of the previous launch:
  * safety   — a word a reader still needs is never overwritten: when block B polls parity t & 1 of neighbour A it can only
               ever see tag t - 2 (not yet published) or tag t, never t + 2;
  * liveness — every block finishes all n steps (no deadlock), whatever the interleaving;
  * values   — the value read with tag t is the one A published for step t.
Pure Python, no GPU."""
import random

import pytest


def neighbours(nbx, nby, b):
    bx, by = b % nbx, b // nbx
    out = []
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if (dx or dy) and 0 <= bx + dx < nbx and 0 <= by + dy < nby:
                out.append((by + dy) * nbx + bx + dx)
    return out


def simulate(nbx, nby, n, epoch0, seed, starve, mailbox=None):
    """Blocks are state machines with the kernel's program order per step t = 1..n:
         publish(t)  ->  poll every neighbour for tag epoch0 + t (one neighbour per scheduler turn)  ->  next step
       (the last step publishes nothing: its result goes to the state planes).  Returns the mailbox for the next launch."""
    rnd = random.Random(seed)
    nb = nbx * nby
    nbrs = [neighbours(nbx, nby, b) for b in range(nb)]
    # mailbox[b][parity] = (tag, value); zero-initialised like the device buffer
    mailbox = mailbox or [[(0, None), (0, None)] for _ in range(nb)]
    step = [1] * nb                 # step being worked on
    published = [False] * nb        # rim of the current step published?
    pending = [[] for _ in range(nb)]
    slow = set(rnd.sample(range(nb), k=max(1, nb // 3))) if starve else set()
    guard = 0
    while any(s <= n for s in step):
        guard += 1
        assert guard < 200000, "deadlock: no block can make progress"
        live = [b for b in range(nb) if step[b] <= n]
        rnd.shuffle(live)
        if starve and rnd.random() < 0.9:
            live = [b for b in live if b not in slow] + [b for b in live if b in slow]
        progressed = False
        for b in live:
            t = step[b]
            tag = epoch0 + t
            if t == n:                                    # last step: no exchange
                step[b] = n + 1
                progressed = True
                break
            if not published[b]:
                old = mailbox[b][tag & 1][0]
                assert old in (0, tag - 2) or old < epoch0 + 1, f"block {b} overwrites tag {old} with {tag}"
                mailbox[b][tag & 1] = (tag, (b, t))
                published[b] = True
                pending[b] = list(nbrs[b])
                progressed = True
                break
            if pending[b]:
                a = pending[b][-1]
                seen_tag, val = mailbox[a][tag & 1]
                # safety: the reader may find the word not yet written (an older tag), never a newer one
                assert seen_tag <= tag, f"block {b} polling step {tag} of block {a} found tag {seen_tag}: overwritten"
                if seen_tag == tag:
                    assert val == (a, t)
                    pending[b].pop()
                    progressed = True
                    break
                continue                                   # not there yet: this block cannot move, try another
            step[b] = t + 1
            published[b] = False
            progressed = True
            break
        if not progressed:
            # every live block is polling a word that is not there: someone must still be able to publish
            raise AssertionError("deadlock: all blocks are waiting")
    return mailbox


@pytest.mark.parametrize("nbx,nby", [(1, 1), (2, 1), (1, 5), (4, 37 // 4), (3, 3), (4, 6)])
@pytest.mark.parametrize("starve", [False, True])
def test_two_parities_are_enough_and_nothing_deadlocks(nbx, nby, starve):
    for seed in range(12):
        simulate(nbx, nby, n=9, epoch0=0, seed=seed, starve=starve)


def test_step_tags_run_on_across_launches():
    """A second launch (epoch0 = steps run so far) finds the first launch's words in the mailbox: their tags are all
    smaller than its first tag, so nothing stale can be mistaken for fresh data — including when the first launch had an
    odd number of steps and the parities swap roles."""
    for n1 in (1, 2, 3, 10):
        for seed in range(6):
            mb = simulate(3, 4, n=n1, epoch0=0, seed=seed, starve=True)
            assert max(tag for blk in mb for tag, _ in blk) <= n1 - 1 or n1 == 1
            mb = simulate(3, 4, n=7, epoch0=n1, seed=seed + 100, starve=bool(seed & 1), mailbox=mb)
            simulate(3, 4, n=4, epoch0=n1 + 7, seed=seed + 200, starve=True, mailbox=mb)


def test_one_parity_would_not_be_enough():
    """The same protocol with a single mailbox slot per block loses words (a fast block publishes step t + 1 over step t
    before a slow neighbour has read it): the model finds it, i.e. the check above is not vacuous."""
    def run(seed):
        rnd = random.Random(seed)
        nb = 3
        mailbox = [(0, None)] * nb
        step, published, pending = [1] * nb, [False] * nb, [[] for _ in range(nb)]
        nbrs = [[1], [0, 2], [1]]
        for _ in range(20000):
            live = [b for b in range(nb) if step[b] <= 9]
            if not live:
                return "ok"
            b = rnd.choice(live)
            t = step[b]
            if t == 9:
                step[b] = 10
            elif not published[b]:
                mailbox[b] = (t, (b, t)); published[b] = True; pending[b] = list(nbrs[b])
            elif pending[b]:
                seen, _ = mailbox[pending[b][-1]]
                if seen > t:
                    return "overwritten"
                if seen == t:
                    pending[b].pop()
            else:
                step[b] = t + 1; published[b] = False
        return "stuck"
    assert any(run(seed) == "overwritten" for seed in range(50))
