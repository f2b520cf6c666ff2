"""CPU restatements of the index logic the round-2 pair kernels rely on (simplefluid_b200/csrc/sf_pairs.cuh,
sf_kernels.cuh), operation by operation, checked exhaustively or under random inputs:

  * the neighbour-list entry codec (list_entry_fluid / list_entry_wall / entry_*_off): the two shared-memory byte
    offsets a walker needs, packed into 32 bits;
  * walk_list: the software pipeline of the list walkers (rows requested unconditionally inside the column, two register
    sets, tail without a further request) visits the first nF entries exactly once, in order, and never reads a row
    outside the column;
  * phase A of k_density_brick (quad-aligned filter reads, miss bits shifted into place by the funnel shift, range and
    self masks with PTX shift semantics): a window's mask reports exactly the filter hits of the lane's own run, in
    slot order, never the particle itself, whatever the run lengths and alignments of the other lanes of the warp;
  * the counted exact phase of k_density_brick (non-empty windows pooled per lane, sentinel entry, refill predicated on
    an empty mask, exactly nh iterations): hits come out in ascending (window, slot) order, nothing is evaluated twice,
    the pool is never indexed beyond its kPool + 2 entries; and its list position, kept as the bytes left in the
    column (stores exactly the first kmax rows, counts every accepted pair even past the capacity);
  * count_into_cell (warp-aggregated arrival ranks of the counting sort, also run by the integrate kernel): the ranks
    of the particles of a cell are a permutation of 0 .. count-1 whatever the order in which the warps arrive, and the
    scatter they drive is a bijection onto the sorted slots.

These are models of the kernels' control flow, not of their arithmetic; the bit-exactness of the arithmetic is what the
GPU parity tests check."""
import random

import numpy as np
import pytest

K_STAGE_CAP = 3584   # halo particles per staging buffer (kStageCap)
K_TAB = 10000        # last kernel-table index (kTab)
K_POOL = 12          # windows pooled per drain (kPool)


# ---------------------------------------------------------------------------------------------------------------
def entry_fluid(halo, tab):
    return ((tab << 18) + (halo << 4)) & 0xFFFFFFFF


def entry_wall(b, tab):
    return ((tab << 18) | b) & 0xFFFFFFFF


def test_list_entry_codec_roundtrip_and_ranges():
    halo = np.arange(K_STAGE_CAP + 512, dtype=np.uint64)  # up to the static bound of 4096
    halo = halo[halo < 4096]
    for tab in (0, 1, 2499, 9999, K_TAB):
        e = (np.uint64(tab) << np.uint64(18)) + (halo << np.uint64(4))
        assert (e < 2**32).all()
        assert np.array_equal(e & np.uint64(0xFFFF), halo * 16)   # byte offset of the float4 in the staging buffer
        assert np.array_equal(e >> np.uint64(16), np.full_like(halo, 4 * tab))  # byte offset in the table
    # the largest offsets stay inside a staging buffer / the table
    assert (entry_fluid(K_STAGE_CAP - 1, K_TAB) & 0xFFFF) + 16 <= K_STAGE_CAP * 16
    assert (entry_fluid(K_STAGE_CAP - 1, K_TAB) >> 16) + 4 <= (K_TAB + 4) * 4
    for b in (0, 1, 62, 63, 1000):  # wall entries: index of the wall particle, not multiplied
        e = entry_wall(b, 777)
        assert e & 0xFFFF == b and e >> 16 == 4 * 777
    # the decoder of the parity download
    e = entry_fluid(1234, 5678)
    assert (e & 0xFFFF) >> 4 == 1234 and (e >> 16) >> 2 == 5678


# ---------------------------------------------------------------------------------------------------------------
def walk_list_model(column, nF, kmax):
    """walk_list of sf_pairs.cuh: `column` is the list column (kmax rows); returns the entries passed to f, in order,
    and the set of rows that were read."""
    read = set()

    def load4(row):
        out = []
        for r in range(row, row + 4):
            assert 0 <= r < kmax, f"row {r} outside the column of {kmax} rows"
            read.add(r)
            out.append(column[r])
        return out

    c = load4(0)  # the caller's rows 0..3, requested before the count is known (every column has >= 8 rows)
    d = [None] * 4
    out = []
    lp = 0
    k = 0
    sets = [c, d]
    cur = 0
    while True:
        e = sets[cur]
        if k + 4 > nF:  # tail: at most three entries, already in registers
            r = nF - k
            for i in range(r):
                out.append(e[i])
            break
        lp += 4
        if k + 8 <= kmax:  # unconditional inside the column
            sets[cur ^ 1][:] = load4(lp)
        out.extend(e)
        k += 4
        cur ^= 1
    return out, read


def test_walk_list_needs_a_capacity_that_is_a_multiple_of_four():
    """Why sf_set_list_capacity rounds up: with 9 rows the guard `k + 8 <= kmax` never requests row 8."""
    column = list(range(9))
    out, _ = walk_list_model(column, 9, 9)
    assert out != column  # (the model of the hazard; the library never runs with such a capacity)


@pytest.mark.parametrize("kmax", [8, 12, 24, 64, 96, 100])
def test_walk_list_visits_every_entry_once_in_order(kmax):
    column = list(range(1000, 1000 + kmax))
    for nF in range(0, kmax + 1):
        out, read = walk_list_model(column, nF, kmax)
        assert out == column[:nF], (kmax, nF)
        assert max(read) < kmax
        # the pipeline never runs more than two batches ahead of what it consumes
        assert max(read) <= min(kmax - 1, (nF // 4) * 4 + 7)


# ---------------------------------------------------------------------------------------------------------------
def counted_exact_phase_model(window_masks, window_bases, self_slot):
    """Phase A + drain of k_density_brick for ONE lane.  window_masks / window_bases: the hit mask and first halo slot
    of every window of the lane's nine candidate runs, in traversal order (a mask may be 0).  Returns the halo slots
    evaluated, in order."""
    pool = [None] * (K_POOL + 2)
    evaluated = []
    ne = nh = nwin = 0

    def ffs(x):
        return (x & -x).bit_length()

    def drain():
        nonlocal ne, nh, nwin
        if nh:
            pool[ne] = (1, self_slot)  # sentinel
            cur, wb = pool[0]
            ei = 1
            e = pool[1]  # may be stale / None: only used after the window in `cur` is exhausted
            j = wb + ffs(cur) - 1
            cur &= cur - 1
            x = j  # "position" of the current hit
            n = nh
            while True:
                # SF_HIT_STEP: evaluate x, extract the next hit
                evaluated.append(x)
                if cur == 0:
                    assert e is not None and ei + 1 < len(pool), "refill beyond the sentinel"
                    cur, wb = e
                    ei += 1
                    e = pool[ei]
                jn = wb + ffs(cur) - 1
                cur &= cur - 1
                x = jn
                n -= 1
                if n == 0:
                    break
            # what was extracted last and never evaluated is the sentinel's slot, or a hit of... no: exactly the sentinel
            assert x == self_slot or nh == 0
        for i in range(len(pool)):
            pool[i] = pool[i]  # (entries stay: stale ones must never be consumed)
        ne = nwin = 0
        nh = 0

    for mask, base in zip(window_masks, window_bases):
        pool[ne] = (mask, base)  # overwritten by the next window when empty
        ne += 1 if mask else 0
        nh += bin(mask).count("1")
        nwin += 1
        if nwin == K_POOL:
            drain()
    if nwin:
        drain()
    return evaluated


def test_counted_exact_phase_walks_the_hits_in_traversal_order():
    rng = random.Random(5)
    for trial in range(3000):
        nw = rng.choice([1, 2, 9, 10, 11, 12, 13, 18, 24, 25, 40])
        masks, bases = [], []
        base = rng.randrange(0, 50)
        for _ in range(nw):
            density = rng.choice([0.0, 0.05, 0.15, 0.5, 1.0])
            m = 0
            for b in range(32):
                if rng.random() < density:
                    m |= 1 << b
            masks.append(m)
            bases.append(base)
            base += rng.randrange(32, 120)  # windows of later rows lie at higher halo slots
        self_slot = 100000 + trial
        want = [b + i for m, b in zip(masks, bases) for i in range(32) if m >> i & 1]
        got = counted_exact_phase_model(masks, bases, self_slot)
        assert got == want


# ---------------------------------------------------------------------------------------------------------------
M32 = 0xFFFFFFFF


def shl_clamp(a, n):
    """PTX shl.b32: shift amounts above 31 give 0 (n is an unsigned 32-bit register)."""
    n &= M32
    return (a << n) & M32 if n < 32 else 0


def filter_quads_model(miss, addr, nq):
    """filter_quads of sf_pairs.cuh with the half-precision arithmetic replaced by an oracle `miss(slot)`: the MISS bits
    of nq quads starting at halo slot `addr` in the top 4 * nq bits (first quad lowest), zeros below; also returns the
    slots read."""
    mask, read = 0, []
    for q in range(nq):
        nb = 0
        for i in range(4):
            read.append(addr + 4 * q + i)
            nb |= (1 if miss(addr + 4 * q + i) else 0) << (28 + i)
        mask = (mask >> 4) | nb
    return mask, read


def window_masks_model(lanes, hit):
    """Phase A of k_density_brick for one halo row of one warp.  lanes: (jbase, length, self_slot) per lane -- the
    lane's candidate run and its own halo slot; hit(lane, slot): what the conservative filter says.  Returns, per lane,
    the (mask, first slot) of every window, exactly as pushed to the pool, and checks that no quad load starts before
    the lane's quad-aligned run start or runs more than one quad past what the widest lane needs."""
    pre = [jb - (jb & ~3) for jb, _, _ in lanes]
    maxlen = max(ln for _, ln, _ in lanes)
    maxend = max((p + ln if ln else 0) for p, (_, ln, _) in zip(pre, lanes))
    out = [[] for _ in lanes]
    c0 = 0
    while c0 < maxlen:
        quads = (min(maxend - c0, 35) + 3) >> 2  # warp-uniform
        assert 1 <= quads <= 9
        nq_lo = min(quads, 8)
        for li, (jbase, ln, self_slot) in enumerate(lanes):
            a0 = jbase & ~3
            addr = a0 + c0
            miss = lambda slot: not hit(li, slot)  # noqa: E731
            lo, read = filter_quads_model(miss, addr, nq_lo)
            hi = 0
            if quads > 8:
                h, r2 = filter_quads_model(miss, addr + 32, 1)
                hi = h >> 28
                read += r2
            assert min(read) == addr and max(read) < addr + 4 * quads
            lo >>= 4 * (8 - nq_lo)
            mask = ~(((hi << 32) | lo) >> pre[li]) & M32  # ~__funnelshift_r(lo, hi, pre)
            vlen = max(ln - c0, 0)
            mask &= ~shl_clamp(M32, vlen) & M32
            mask &= ~shl_clamp(1, (self_slot - (jbase + c0)) & M32) & M32
            # every slot the mask may report was actually read
            for i in range(32):
                if mask >> i & 1:
                    assert jbase + c0 + i in read
            out[li].append((mask, jbase + c0))
        c0 += 32
    return out


def test_window_masks_report_exactly_the_filter_hits_of_the_lanes_own_run():
    rng = random.Random(11)
    for trial in range(400):
        nl = rng.choice([1, 4, 32])
        lanes = []
        for _ in range(nl):
            ln = rng.choice([0, 1, 3, 17, 26, 31, 32, 33, 35, 36, 40, 56, 64, 65, 100])
            jbase = rng.randrange(0, 3000)
            self_slot = rng.choice([jbase + rng.randrange(0, max(ln, 1)), jbase - 5 if jbase >= 5 else jbase + ln + 7, rng.randrange(0, 3584)])
            lanes.append((jbase, ln, self_slot))
        density = rng.choice([0.0, 0.15, 0.6, 1.0])
        table = {}

        def hit(li, slot):
            return table.setdefault((li, slot), rng.random() < density)

        got = window_masks_model(lanes, hit)
        for li, (jbase, ln, self_slot) in enumerate(lanes):
            slots = [b + i for m, b in got[li] for i in range(32) if m >> i & 1]
            want = [sl for sl in range(jbase, jbase + ln) if sl != self_slot and hit(li, sl)]
            assert slots == want, (trial, li, lanes[li])


# ---------------------------------------------------------------------------------------------------------------
def list_countdown_model(passes, kmax, column_base):
    """The list position of the density pass's exact phase (SF_HIT_STEP): ko = bytes LEFT in the column as a 32-bit
    register, the store goes to (column end - ko) when the pair passes and (int32)ko > 0, and ko drops by 128 (one
    row) per accepted pair -- also beyond the capacity, where it wraps below zero.  Returns (stored addresses,
    accepted pairs recovered from ko, address of the next row)."""
    M = 0xFFFFFFFF
    kmax_bytes = kmax * 128
    ko = kmax_bytes
    lpe = column_base + kmax_bytes
    stored = []
    for p in passes:
        signed = ko - (1 << 32) if ko & 0x80000000 else ko
        if p and signed > 0:
            stored.append(lpe - ko)
        if p:
            ko = (ko + ((-128) & M)) & M
    words = ((kmax_bytes - ko) & M) >> 2
    return stored, words // 32, column_base + 4 * words


@pytest.mark.parametrize("kmax", [8, 28, 64, 16380])
def test_list_countdown_stores_the_first_kmax_rows_and_counts_every_pair(kmax):
    rng = random.Random(kmax)
    base = 0x7F0000001000
    for trial in range(200):
        n = rng.choice([0, 1, kmax - 1, kmax, kmax + 1, 2 * kmax + 3, rng.randrange(0, 3 * kmax)])
        passes = [rng.random() < 0.8 for _ in range(n)]
        stored, accepted, next_row = list_countdown_model(passes, kmax, base)
        a = sum(passes)
        assert accepted == a                                            # the count survives the overflow (fits = k <= kmax)
        assert stored == [base + 128 * r for r in range(min(a, kmax))]  # rows 0 .. kmax-1 in order, nothing past the column
        assert next_row == base + 128 * a                               # where the wall rows continue (only used when a < kmax)


# ---------------------------------------------------------------------------------------------------------------
def count_into_cell_model(keys_by_warp, order, ncells):
    """k_hash_count / the fused part of k_visc_brick: warps arrive in `order`; inside a warp one atomicAdd per distinct
    key (leader = lowest lane with that key) and rank = base + number of lower lanes with the same key."""
    cell_cnt = np.zeros(ncells, dtype=np.int64)
    ranks = {}
    for w in order:
        keys = keys_by_warp[w]
        base_of = {}
        for lane, key in enumerate(keys):
            if key is None:  # inactive lane: invalid key, no atomic
                continue
            if key not in base_of:
                peers = sum(1 for k in keys if k == key)
                base_of[key] = int(cell_cnt[key])
                cell_cnt[key] += peers
            lower = sum(1 for k in keys[:lane] if k == key)
            ranks[(w, lane)] = base_of[key] + lower
    return cell_cnt, ranks


def test_arrival_ranks_are_a_permutation_per_cell_for_any_warp_order():
    rng = random.Random(9)
    for _ in range(200):
        ncells = rng.choice([1, 3, 17, 64])
        nwarps = rng.randrange(1, 12)
        keys_by_warp = [[(rng.randrange(ncells) if rng.random() < 0.9 else None) for _ in range(32)] for _ in range(nwarps)]
        order = list(range(nwarps))
        rng.shuffle(order)
        cnt, ranks = count_into_cell_model(keys_by_warp, order, ncells)
        begin = np.concatenate(([0], np.cumsum(cnt)[:-1]))  # the exclusive scan of k_cell_scan_*
        per_cell = {}
        slots = []
        for (w, lane), r in ranks.items():
            key = keys_by_warp[w][lane]
            per_cell.setdefault(key, []).append(r)
            slots.append(int(begin[key]) + r)  # k_count_scatter: dst = cellTab[key].x + rank
        for key, rs in per_cell.items():
            assert sorted(rs) == list(range(int(cnt[key])))
        n = sum(1 for ks in keys_by_warp for k in ks if k is not None)
        assert sorted(slots) == list(range(n))  # a bijection onto the sorted slots
