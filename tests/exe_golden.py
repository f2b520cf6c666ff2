"""Loader for tests/golden/exe_*.npz -- outputs of the reference's own compiled SPH step (tests/golden/make_exe_golden.py)
-- and a generic driver that replays a case on any solver with the oracle's method names."""
import ast
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(f)[4:-4] for f in glob.glob(os.path.join(GOLDEN, "exe_*.npz")))


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, f"exe_{name}.npz")))
    g["scene"], g["resolution"], g["seed"] = str(g["scene"]), float(g["resolution"]), int(g["seed"])
    g["overrides"] = ast.literal_eval(str(g["overrides"]))
    return g


def replay(g, advance, fields, max_steps=None):
    """advance() -> dt of one substep; fields() -> dict(cell, rho, acc, x, v) of the state after it.  Asserts the dt
    sequence and every stored field bit for bit.  Returns the number of substeps compared."""
    steps = len(g["dts"]) if max_steps is None else min(max_steps, len(g["dts"]))
    kept = set(int(k) for k in g["steps_kept"])
    for k in range(steps):
        dt = advance()
        assert np.float32(dt) == g["dts"][k], f"dt of substep {k}: {dt!r} != reference {g['dts'][k]!r}"
        if k in kept:
            f = fields()
            for name in ("cell", "rho", "acc", "x", "v"):
                if name in f and f[name] is not None:
                    a, b = np.asarray(f[name]), g[f"{name}{k}"]
                    assert a.shape == b.shape and a.tobytes() == b.tobytes(), \
                        f"{name} after substep {k} differs from the reference binary's output ({int((a != b).sum())} values)"
    return steps
