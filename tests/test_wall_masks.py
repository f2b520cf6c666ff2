"""The candidate masks of the wall lists (simplefluid_b200/csrc/sf_host.cpp: wall_candidate_masks / wall_subcell) that let
the density pass skip wall particles that cannot be within h.  The reference tests every particle of a wall list
(EXE@0x1400179c0); skipping is only legal when it is conservative: every wall particle the exact fp32 test
accepts for a position must be marked in the mask of the sub-cell that position maps to.  Checked here by brute force
with the kernel's own arithmetic restated in numpy float32 (wall_of / wall_shift / dist2 of sf_kernels.cuh), for the
generated walls at several resolutions and for arbitrary user-supplied wall lists."""
import numpy as np
import pytest

f32 = np.float32


def near_wall_positions(p, wall, n, rng, h):
    """Positions inside the box within h of `wall`, including the clamp positions bound +- r the integrator produces,
    positions exactly on the box faces and on sub-cell borders."""
    A = wall // 2
    x = rng.uniform(-1.0, 1.0, (n, 3))
    dn = rng.uniform(0.0, 1.0, n) * h
    dn[: n // 8] = p.particleRadius                  # resting on the wall (A.14 clamp)
    dn[n // 8: n // 6] = 0.0                          # on the face itself
    dn[n // 6: n // 5] = np.round(rng.uniform(0, 4, n // 5 - n // 6)) * h / 4  # sub-cell borders
    x[:, A] = (p.boxMax[A] - dn) if wall & 1 else (p.boxMin[A] + dn)
    k = n // 5
    x[k: 2 * k, (A + 1) % 3] = np.round(x[k: 2 * k, (A + 1) % 3] / (h / 4)) * (h / 4)  # tangential sub-cell borders
    return np.clip(x, -1.0, 1.0).astype(f32)


def exact_hits(p, wall, walls_xyz, x):
    """The kernel's test for one position: wall_of, wall_shift, d2 = (dx*dx + dy*dy) + dz*dz <= radius2, all fp32."""
    A = wall // 2
    h = f32(p.kernelRadius)
    lo, hi = f32(h + f32(p.boxMin[A])), f32(f32(p.boxMax[A]) - h)
    near = (lo > x[A]) if not (wall & 1) else (x[A] > hi)
    if not near:
        return None
    xs = x.copy()
    for d in range(3):
        if d != A:
            xs[d] = f32(x[d] - f32(h * np.floor(f32(x[d] / h))))
    d = (walls_xyz - xs).astype(f32)
    d2 = (f32(1) * d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32) + (d[:, 2] * d[:, 2]).astype(f32)
    return np.nonzero(f32(p.kernelRadiusSqr) >= d2.astype(f32))[0]


def mask_bits(masks, entry):
    row = masks[entry]
    return {w * 32 + b for w in range(len(row)) for b in range(32) if row[w] >> np.uint32(b) & np.uint32(1)}


@pytest.mark.parametrize("res", [24.0, 61.0, 203.0])
def test_generated_walls_every_exact_hit_is_a_candidate(sf, res):
    from simplefluid_b200 import binding as B
    p = sf.default_params(res, "Dambreak")
    rng = np.random.default_rng(int(res))
    total_hits = total_cand = 0
    for wall in range(6):
        w = B.boundary_generate(p, 0, wall)
        masks = B.wall_candidate_masks(p, wall, w)
        assert masks.shape == (B.WALL_SUBCELLS + 1, (len(w) + 31) // 32)
        assert mask_bits(masks, B.WALL_SUBCELLS) == set(range(len(w)))  # the catch-all entry marks every particle
        bits = [mask_bits(masks, e) for e in range(B.WALL_SUBCELLS)]
        for x in near_wall_positions(p, wall, 1500, rng, p.kernelRadius):
            hits = exact_hits(p, wall, w, x)
            if hits is None:
                continue
            e = B.wall_subcell(p, wall, x)
            assert e < B.WALL_SUBCELLS
            assert set(hits.tolist()) <= bits[e], (wall, x, e)
            total_hits += len(hits)
            total_cand += len(bits[e])
    assert total_hits > 1000           # the test positions do meet the walls
    assert total_cand < 6 * total_hits  # ... and the masks do cull: a few candidates per hit, not 243 / 9


def test_arbitrary_wall_lists_and_positions_beyond_the_wall_plane(sf):
    from simplefluid_b200 import binding as B
    p = sf.default_params(50.0, "CubeDrop")
    h = p.kernelRadius
    rng = np.random.default_rng(7)
    for wall in range(6):
        A = wall // 2
        n = int(rng.integers(1, 700))
        w = rng.uniform(-0.5 * h, 1.5 * h, (n, 3))
        depth = rng.uniform(-0.2 * h, 1.2 * h, n)  # some even inside the box
        w[:, A] = (p.boxMax[A] + depth) if wall & 1 else (p.boxMin[A] - depth)
        w = w.astype(f32)
        masks = B.wall_candidate_masks(p, wall, w)
        bits = [mask_bits(masks, e) for e in range(B.WALL_SUBCELLS + 1)]
        for x in near_wall_positions(p, wall, 600, rng, h):
            hits = exact_hits(p, wall, w, x)
            if hits is not None:
                assert set(hits.tolist()) <= bits[B.wall_subcell(p, wall, x)]
        # a position beyond the wall plane (an upload in the partly outside last cell layer) takes the catch-all entry
        x = np.zeros(3, f32)
        x[A] = f32(p.boxMax[A] + 0.3 * h) if wall & 1 else f32(p.boxMin[A] - 0.3 * h)
        assert B.wall_subcell(p, wall, x) == B.WALL_SUBCELLS
    with pytest.raises(B.SFError):
        B.wall_candidate_masks(p, 6, np.zeros((3, 3), f32))
