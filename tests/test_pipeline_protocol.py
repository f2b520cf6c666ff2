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


def test_counting_sort_and_reorder_give_the_reference_cell_lists():
    """Restatement of the sort pipeline (k_hash_count -> scan -> k_count_scatter -> k_reorder): whatever order the
    atomics hand out inside a cell, re-ranking by original id yields the reference's cell lists -- cells in key order,
    ascending particle id inside a cell (collectParticlesToCells pushes ids in ascending order, EXE@0x140016890)."""
    import numpy as np
    rng = np.random.default_rng(11)
    for _ in range(20):
        n, ncells = int(rng.integers(1, 4000)), int(rng.integers(1, 300))
        ids = rng.permutation(n).astype(np.uint32)          # A order: last substep's sorted order, arbitrary ids
        keys = rng.integers(0, ncells, size=n).astype(np.uint32)
        dead = rng.random(n) < 0.1                           # slab mode: last substep's ghosts
        count = np.zeros(ncells, np.uint32)
        rank = np.zeros(n, np.uint32)
        for i in rng.permutation(n):                         # arrival order of the atomics: arbitrary
            if not dead[i]:
                rank[i] = count[keys[i]]
                count[keys[i]] += 1
        begin = np.concatenate(([0], np.cumsum(count)[:-1])).astype(np.uint32)
        live = int(count.sum())
        slot_of = np.full(live, -1, np.int64)                # k_count_scatter
        for i in range(n):
            if not dead[i]:
                slot_of[begin[keys[i]] + rank[i]] = i
        assert (slot_of >= 0).all()
        out = np.full(live, 0xFFFFFFFF, np.uint32)           # k_reorder: rank by id inside the cell
        for p in range(live):
            src = slot_of[p]
            c = keys[src]
            members = slot_of[begin[c]:begin[c] + count[c]]
            out[begin[c] + np.count_nonzero(ids[members] < ids[src])] = ids[src]
        alive = np.flatnonzero(~dead)
        want = ids[alive][np.lexsort((ids[alive], keys[alive]))]
        assert np.array_equal(out, want)
