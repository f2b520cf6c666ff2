"""Lane occupancy of the exact phase (phase B) of k_density_brick, estimated on the CPU from the oracle's neighbour
sets: per warp of 32 consecutive sorted particles, the number of phase-B iterations is the maximum over its lanes of
the hits they have to process between two points where the warp re-converges.  Compares the current scheme (hits of
one halo row at a time) with pooling the hits of a dz-plane (3 rows) or of all 9 rows.  Analysis aid for DESIGN.md.

    python tests/tools/phaseb_occupancy.py [scene] [res] [substeps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402  (test infrastructure; this tool is an analysis aid, not product code)


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "Dambreak"
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    p = ob.default_params(res, scene)
    pos = ob.scene(p)
    orc = ob.Oracle(p, pos, boundary_seed=0)
    for _ in range(steps):
        orc.advance()
    cnt, ids = orc.neighbors()
    orc.advance()  # bins the particles: cell_index() is that of the substep just taken
    key = orc.cell_index().astype(np.int64)
    n = len(cnt)
    nx = ny = res
    cy, cz = (key // nx) % ny, key // (nx * ny)
    off = np.concatenate(([0], np.cumsum(cnt)))
    owner = np.repeat(np.arange(n), cnt)
    row = (cz[ids] - cz[owner] + 1) * 3 + (cy[ids] - cy[owner] + 1)  # 0..8, reference order: dz outer, dy inner
    ok = (row >= 0) & (row < 9)
    hits = np.zeros((n, 9), np.int64)
    np.add.at(hits, (owner[ok], row[ok]), 1)
    order = np.lexsort((np.arange(n), key))
    h = hits[order]
    pad = (-n) % 32
    h = np.concatenate((h, np.zeros((pad, 9), np.int64))).reshape(-1, 32, 9)
    total = h.sum()
    per_row = h.max(axis=1).sum()
    per_plane = h.reshape(-1, 32, 3, 3).sum(axis=3).max(axis=1).sum()
    all_rows = h.sum(axis=2).max(axis=1).sum()
    warps = h.shape[0]
    print(f"{scene} res {res} after {steps} substeps: N={n}, {total / n:.1f} neighbours per particle")
    for name, iters in (("one row at a time (current)", per_row), ("one dz-plane (3 rows) pooled", per_plane), ("all 9 rows pooled", all_rows)):
        print(f"  {name:32s} {iters / warps:6.1f} iterations per warp, lane occupancy {total / (iters * 32) * 100:5.1f} %")


if __name__ == "__main__":
    main()
