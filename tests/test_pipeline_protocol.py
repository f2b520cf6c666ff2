"""Model check of the producer/consumer protocol of the pair kernels (simplefluid_b200/csrc/sf_pairs.cuh:
producer_loop + the consumer loops): NBUF staging buffers, NBUF + 1 meta slots, one `full` (1 arrival) and one `empty`
(one arrival per consumer warp) mbarrier per meta slot, waits by phase parity.  The index arithmetic of the kernels is
restated here (slot = brick % (NBUF + 1), buffer = brick % NBUF, the producer waits for `empty` of brick - NBUF, which
lives in slot_next(slot)) and run under random interleavings of the producer and the consumer warps; the invariants:
nobody reads a meta slot or a staging buffer that holds another brick, the producer never overwrites one that is
still in use, parities never alias, everybody terminates."""
import random

import pytest


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test_wait(self, parity):  # mbarrier.test_wait.parity: has the phase of that parity completed?
        return (self.phase & 1) != parity


def simulate(nbuf, nbricks, warps, seed):
    rng = random.Random(seed)
    slots = nbuf + 1
    full = [Barrier(1) for _ in range(slots)]
    empty = [Barrier(warps) for _ in range(slots)]
    meta = [None] * slots    # brick whose tables the slot holds (-1: end marker)
    stage = [None] * nbuf    # brick whose halo the buffer holds
    holding = {}             # consumer -> (slot, buf, brick) while it works on a brick

    def producer():
        pe = [0] * slots
        it, slot, buf = 0, 0, 0
        while True:
            # brick_prepare into meta[slot]: no consumer may still be reading that slot
            assert all(h[0] != slot for h in holding.values()), "meta slot overwritten while in use"
            yield
            if it >= nbricks:
                meta[slot] = -1
                full[slot].arrive()
                return
            meta[slot] = it
            yield
            if it >= nbuf:
                s2 = (slot + 1) % slots  # slot of brick it - nbuf
                assert (it - nbuf) % slots == s2
                while not empty[s2].test_wait(pe[s2]):
                    yield
                pe[s2] ^= 1
            # brick_issue into stage[buf]
            assert all(h[1] != buf for h in holding.values()), "staging buffer overwritten while in use"
            assert stage[buf] is None or stage[buf] == it - nbuf
            stage[buf] = it
            yield
            full[slot].arrive()  # TMA complete_tx
            it, slot, buf = it + 1, (slot + 1) % slots, (buf + 1) % nbuf

    def consumer(cid):
        ph = [0] * slots
        it, slot, buf = 0, 0, 0
        while True:
            while not full[slot].test_wait(ph[slot]):
                yield
            ph[slot] ^= 1
            if meta[slot] == -1:
                return
            assert meta[slot] == it, f"consumer {cid} expected brick {it} in slot {slot}, found {meta[slot]}"
            assert stage[buf] == it, f"consumer {cid} expected brick {it} in buffer {buf}, found {stage[buf]}"
            holding[cid] = (slot, buf, it)
            for _ in range(rng.randint(0, 3)):  # groups of this brick
                yield
                assert meta[slot] == it and stage[buf] == it
            del holding[cid]
            empty[slot].arrive()
            it, slot, buf = it + 1, (slot + 1) % slots, (buf + 1) % nbuf
            yield

    actors = [producer()] + [consumer(c) for c in range(warps)]
    alive = list(range(len(actors)))
    steps = 0
    while alive:
        i = rng.choice(alive)
        try:
            next(actors[i])
        except StopIteration:
            alive.remove(i)
        steps += 1
        assert steps < 2_000_000, "deadlock or livelock"
    return steps


@pytest.mark.parametrize("nbuf", [2, 3])
def test_protocol_random_interleavings(nbuf):
    for seed in range(60):
        rng = random.Random(1000 + seed)
        simulate(nbuf, nbricks=rng.choice([0, 1, 2, 3, 4, 5, 7, 12, 40]), warps=rng.choice([1, 2, 5, 31]), seed=seed)


def test_protocol_needs_the_extra_meta_slot():
    """With as many meta slots as buffers the early preparation would overwrite tables in use: the model must notice
    (guards the test itself against being vacuous)."""
    def broken(seed):
        # same protocol, but the producer prepares into slot = brick % nbuf (no spare slot)
        rng = random.Random(seed)
        nbuf, warps, nbricks = 2, 3, 12
        meta, holding = [None] * nbuf, {}
        full = [Barrier(1) for _ in range(nbuf)]
        empty = [Barrier(warps) for _ in range(nbuf)]

        def producer():
            pe = [0] * nbuf
            for it in range(nbricks + 1):
                slot = it % nbuf
                assert all(h != slot for h in holding.values()), "meta slot overwritten while in use"
                meta[slot] = it if it < nbricks else -1
                yield
                if it >= nbuf and it < nbricks:
                    while not empty[slot].test_wait(pe[slot]):
                        yield
                    pe[slot] ^= 1
                full[slot].arrive()

        def consumer(cid):
            ph = [0] * nbuf
            it = 0
            while True:
                slot = it % nbuf
                while not full[slot].test_wait(ph[slot]):
                    yield
                ph[slot] ^= 1
                if meta[slot] == -1:
                    return
                holding[cid] = slot
                for _ in range(rng.randint(1, 3)):
                    yield
                del holding[cid]
                empty[slot].arrive()
                it += 1
                yield

        actors = [producer()] + [consumer(c) for c in range(warps)]
        alive = list(range(len(actors)))
        for _ in range(100_000):
            if not alive:
                return
            i = rng.choice(alive)
            try:
                next(actors[i])
            except StopIteration:
                alive.remove(i)

    with pytest.raises(AssertionError):
        for seed in range(200):
            broken(seed)
