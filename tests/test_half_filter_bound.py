"""CPU check of the numerics behind k_density_brick's half-precision candidate filter (simplefluid_b200/csrc/
sf_pairs.cuh): the filter may let non-neighbours through (phase B re-applies the exact fp32 predicate) but must never
drop a pair whose exact, separately rounded fp32 distance satisfies d2 <= radius2.  The kernel's arithmetic is
restated here operation by operation with numpy (float16 add/mul are correctly rounded; a fused multiply-add in half
precision is emulated by one rounding of the exact float64 result) and hammered with pairs at and around the
boundary, for the grid resolutions of every BASELINE.json config.  Also covered: the bit tricks that assemble the
hit mask of a 32-slot window and the window's range mask."""
import numpy as np
import pytest

HX, HY, HZ = 10, 6, 6  # halo of an 8 x 4 x 4 brick, in cells
F32, F16 = np.float32, np.float16


def f32(x):
    return np.asarray(x, dtype=F32)


def half_fma(a, b, c):
    """fp16 fma: exact in float64 (11-bit significands), one rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F16)


def half_round_up(x32):
    """__float2half_ru of a positive float32 scalar."""
    h = F16(x32)
    return h if F32(h) >= x32 else np.nextafter(h, F16(np.inf))


def filter_passes(xq, xp, centre, h):
    """The filter's verdict for candidate positions xq against own positions xp (float32 [n, 3])."""
    invh = F32(1.0) / F32(h)
    off = f32(-centre * invh)  # ox, oy, oz
    # u = fma(x, invh, o) in fp32, then round to nearest half (producer's conversion)
    uq = (xq.astype(np.float64) * np.float64(invh) + off.astype(np.float64)).astype(F32).astype(F16)
    up = (xp.astype(np.float64) * np.float64(invh) + off.astype(np.float64)).astype(F32).astype(F16)
    d = uq - up  # HADD2, correctly rounded
    radius2 = F32(h) * F32(h)
    thr = half_round_up(F32(F32(F32(radius2 * invh) * invh) * F32(1.0135)))
    # t = ((thr - dx^2) - dy^2) - dz^2 as three half-precision fused multiply-adds; kept when the sign bit of t is clear
    t = half_fma(-d[:, 0], d[:, 0], np.full(len(d), thr, dtype=F16))
    t = half_fma(-d[:, 1], d[:, 1], t)
    t = half_fma(-d[:, 2], d[:, 2], t)
    return ~np.signbit(t), thr


def exact_in_range(xq, xp, h):
    """The reference's predicate: d2 = (dx*dx + dy*dy) + dz*dz in separately rounded fp32, d2 <= h*h."""
    d = xq - xp
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return d2 <= F32(h) * F32(h)


@pytest.mark.parametrize("res", [8, 24, 61, 100, 161, 203, 313, 404, 1000])
def test_filter_never_drops_a_neighbour(res):
    rng = np.random.default_rng(res)
    h = F32(2.0) / F32(res)
    n = 400_000
    # a brick anywhere in the box: halo origin in cells, physical centre as the producer computes it
    cell0 = rng.integers(-1, max(res - HX + 2, 0) + 1, size=3)
    bmin = F32(-1.0)
    centre = f32([bmin + h * F32(cell0[0] + HX // 2), bmin + h * F32(cell0[1] + HY // 2), bmin + h * F32(cell0[2] + HZ // 2)])
    lo = f32(bmin + h * f32(cell0))
    span = f32(h * f32([HX, HY, HZ]))
    xp = f32(lo + span * rng.random((n, 3), dtype=F32))
    # candidates: directions uniform on the sphere; distances concentrated on the boundary band, plus exact duplicates
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    r = np.where(rng.random(n) < 0.8, 1.0 - 3e-3 * rng.random(n), rng.random(n)) * float(h)
    xq = f32(xp + f32(v * r[:, None]))
    xq[:100] = xp[:100]
    # keep both ends inside the halo box (where the kernel's |u| <= 5 / 3 assumption holds)
    inside = np.all((xq >= lo) & (xq <= lo + span), axis=1)
    xq, xp = xq[inside], xp[inside]
    near = exact_in_range(xq, xp, h)
    assert near.sum() > 100_000
    passed, thr = filter_passes(xq, xp, centre, h)
    assert np.all(passed[near]), f"filter dropped {np.count_nonzero(near & ~passed)} true neighbours (threshold {thr})"
    # and it is a filter: far candidates are rejected
    far = f32(xp + f32(v[inside] * (1.2 * float(h))))
    ok_far = np.all((far >= lo) & (far <= lo + span), axis=1)
    passed_far, _ = filter_passes(far[ok_far], xp[ok_far], centre, h)
    assert not passed_far.any()


def test_worst_case_corner_of_the_halo():
    """Both ends at the far corner of the halo (|u| close to 5, 3, 3: the coarsest half-precision grid), distance on
    the boundary, every axis orientation."""
    rng = np.random.default_rng(7)
    for res in (24, 203, 404):
        h = F32(2.0) / F32(res)
        bmin = F32(-1.0)
        cell0 = np.array([3, 2, 1])
        centre = f32([bmin + h * F32(cell0[0] + 5), bmin + h * F32(cell0[1] + 3), bmin + h * F32(cell0[2] + 3)])
        n = 200_000
        corner = f32(bmin + h * f32(cell0 + np.array([HX, HY, HZ])))
        xp = f32(corner - h * f32(rng.random((n, 3)) * 0.9))
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v = -np.abs(v)  # towards the inside of the halo
        r = (1.0 - 1e-3 * rng.random(n)) * float(h)
        xq = f32(xp + f32(v * r[:, None]))
        near = exact_in_range(xq, xp, h)
        passed, thr = filter_passes(xq, xp, centre, h)
        assert near.sum() > 10_000 and np.all(passed[near])


def test_miss_nibble_and_window_mask_assembly():
    """prmt (sign-replicated bytes 1 and 3 of both registers) + LOP3 + IMAD of filter_quads: the sign bits of four
    half-precision t values -> bits 28..31 (miss bits), shifted in quad by quad; the caller inverts them."""
    def prmt_signs(ta, tb):
        b = [(ta >> 8) & 0xFF, (ta >> 24) & 0xFF, (tb >> 8) & 0xFF, (tb >> 24) & 0xFF]
        return sum((0xFF if x & 0x80 else 0) << (8 * i) for i, x in enumerate(b))

    def nibble(ta, tb):
        return ((prmt_signs(ta, tb) & 0x08040201) * 0x10101010) & 0xF0000000

    rng = np.random.default_rng(0)
    for bits in range(16):
        for _ in range(20):
            junk = [int(rng.integers(0, 0x8000)) for _ in range(4)]  # any magnitude bits
            h = [junk[i] | (0x8000 if bits >> i & 1 else 0) for i in range(4)]
            assert nibble(h[0] | h[1] << 16, h[2] | h[3] << 16) >> 28 == bits
    for nq in range(1, 9):
        for _ in range(50):
            miss = rng.integers(0, 2, size=4 * nq)
            mask = 0
            for q in range(nq):
                b = [0x8000 if v else 0x3c00 for v in miss[4 * q:4 * q + 4]]
                mask = ((mask >> 4) | nibble(b[0] | b[1] << 16, b[2] | b[3] << 16)) & 0xFFFFFFFF
            mask >>= 4 * (8 - nq)
            assert mask == sum(int(v) << i for i, v in enumerate(miss))


def test_window_mask_of_the_pooled_exact_phase():
    """The hit mask of one 32-slot window as k_density_brick assembles it: miss bits of the quads read from the
    quad-aligned start a0 + c0 (lo: 8 quads, hi: the ninth), funnel-shifted by pre = jbase - a0, inverted, restricted
    to the lane's own run [0, len - c0) and cleared at the particle's own slot; shifts clamp at 32 (PTX shl)."""
    def shl_clamp(a, n):
        return (a << n) & 0xFFFFFFFF if 0 <= n < 32 else 0

    rng = np.random.default_rng(3)
    for _ in range(4000):
        jbase = int(rng.integers(0, 3000))
        length = int(rng.choice([0, 1, 3, 17, 24, 31, 32, 33, 40, 56, 57, 63, 64, 90]))
        maxend = max((jbase & 3) + length, int(rng.integers(0, 100)))  # some other lane may need more quads
        self_slot = jbase + int(rng.integers(-5, length + 5))
        hits = rng.integers(0, 2, size=jbase + 200)  # filter verdict per halo slot
        a0, pre = jbase & ~3, jbase & 3
        c0 = 0
        while c0 < max(length, 1) and length:
            quads = (min(maxend - c0, 35) + 3) >> 2
            assert 1 <= quads <= 9
            nq_lo = min(quads, 8)
            lo = sum((1 - int(hits[a0 + c0 + i])) << i for i in range(4 * nq_lo))            # after lo >>= 4 * (8 - nqLo)
            hi = sum((1 - int(hits[a0 + c0 + 32 + i])) << i for i in range(4)) if quads > 8 else 0
            mask = ~(((hi << 32 | lo) >> pre) & 0xFFFFFFFF) & 0xFFFFFFFF
            vlen = max(length - c0, 0)
            mask &= ~shl_clamp(0xFFFFFFFF, vlen) & 0xFFFFFFFF
            ts = (self_slot - (jbase + c0)) & 0xFFFFFFFF
            mask &= ~shl_clamp(1, ts if ts < 2**31 else 99) & 0xFFFFFFFF
            want = sum(1 << i for i in range(32)
                       if c0 + i < length and hits[jbase + c0 + i] and jbase + c0 + i != self_slot)
            assert mask == want, (jbase, length, maxend, c0)
            c0 += 32
