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
    d2 = d[:, 0] * d[:, 0]  # HMUL2
    d2 = half_fma(d[:, 1], d[:, 1], d2)
    d2 = half_fma(d[:, 2], d[:, 2], d2)
    radius2 = F32(h) * F32(h)
    thr = half_round_up(F32(F32(F32(radius2 * invh) * invh) * F32(1.0135)))
    return d2 <= thr, thr


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


def test_hit_nibble_and_window_mask_assembly():
    """PRMT + LOP3 + IMAD of filter_quads: four 0xffff / 0 half-masks -> bits 28..31, shifted in quad by quad."""
    def nibble(ma, mb):
        tq = (ma & 0xFF) | (((ma >> 16) & 0xFF) << 8) | ((mb & 0xFF) << 16) | (((mb >> 16) & 0xFF) << 24)  # __byte_perm(ma, mb, 0x6420)
        return ((tq & 0x08040201) * 0x10101010) & 0xF0000000

    for bits in range(16):
        ma = (0xFFFF if bits & 1 else 0) | (0xFFFF0000 if bits & 2 else 0)
        mb = (0xFFFF if bits & 4 else 0) | (0xFFFF0000 if bits & 8 else 0)
        assert nibble(ma, mb) >> 28 == bits
    rng = np.random.default_rng(1)
    for nq in range(1, 9):
        for _ in range(50):
            hits = rng.integers(0, 2, size=4 * nq)
            mask = 0
            for q in range(nq):
                b = hits[4 * q:4 * q + 4]
                ma = (0xFFFF if b[0] else 0) | (0xFFFF0000 if b[1] else 0)
                mb = (0xFFFF if b[2] else 0) | (0xFFFF0000 if b[3] else 0)
                mask = ((mask >> 4) | nibble(ma, mb)) & 0xFFFFFFFF
            mask >>= 4 * (8 - nq)
            assert mask == sum(int(v) << i for i, v in enumerate(hits))


def test_window_range_mask():
    """Bits of a 32-slot window that belong to the lane's own run [pre, pre + len) of the aligned window, per chunk."""
    def range_mask(pre, length, q0):
        lo = pre - 4 * q0
        hi = lo + length
        mhi = 0xFFFFFFFF if hi >= 32 else (0 if hi <= 0 else (1 << hi) - 1)
        mlo = 0xFFFFFFFF if lo <= 0 else (0 if lo >= 32 else ~((1 << lo) - 1) & 0xFFFFFFFF)
        return mhi & mlo

    for pre in range(4):
        for length in (0, 1, 5, 29, 32, 33, 61, 64, 100):
            nq = (pre + length + 3) // 4 if length else 0
            for q0 in range(0, max(nq, 1) + 8, 8):
                want = sum(1 << i for i in range(32) if pre <= 4 * q0 + i < pre + length)
                assert range_mask(pre, length, q0) == want
