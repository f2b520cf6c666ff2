"""Verbose GPU-vs-oracle comparison (development aid; the gated checks live in tests/)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
import simplefluid_b200 as sf  # noqa: E402


def cmp(name, a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    exact = np.array_equal(a, b)
    if a.dtype.kind == "f":
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        scale = np.maximum(np.abs(b).astype(np.float64), 1e-30)
        print(f"  {name:12s} exact={exact}  max_abs={d.max():.3e}  max_rel={np.max(d / scale):.3e}  n_diff={(a != b).sum()}")
    else:
        print(f"  {name:12s} exact={exact}  n_diff={(a != b).sum()}")
    return exact


def check(scene, res, steps, **over):
    print(f"== {scene} res {res} steps {steps} {over}")
    p = sf.default_params(res, scene, **over)
    po = ob.default_params(res, scene, **over)
    pos = sf.scene_generate(p)
    print(f"  N = {len(pos)}")
    orc = ob.Oracle(po, pos, boundary_seed=0)
    gpu = sf.SPHSolver(p)
    gpu.setParticles(pos)
    gpu.generateBoundaryParticles(0)
    gpu.setCapture(True)
    gpu.makeReady()
    ok = True
    ocnt, oids = orc.neighbors()
    t = time.time()
    dto = orc.advance()
    t_or = time.time() - t
    dtg = gpu.advanceFrame()
    print(f"  dt oracle {dto!r} gpu {dtg!r}  oracle step {t_or:.3f}s")
    ok &= cmp("cell_index", gpu.cellIndex(), orc.cell_index())
    gcnt, gids = gpu.neighbors()
    ok &= cmp("nbr_count", gcnt, ocnt)
    ok &= cmp("nbr_ids", gids, oids) if len(gids) == len(oids) else False
    ok &= cmp("density", gpu.density(), orc.density())
    ok &= cmp("pressure", gpu.pressure(), orc.pressure())
    ok &= cmp("accel", gpu.accel(), orc.accel())
    ok &= cmp("velocity", gpu.getVelocity(), orc.velocities())
    ok &= cmp("position", gpu.getParticles(), orc.positions())
    for k in range(1, steps):
        dto = orc.advance()
        dtg = gpu.advanceFrame()
        if dto != dtg:
            print(f"  step {k}: dt differs {dto!r} {dtg!r}")
            ok = False
            break
    if steps > 1:
        print(f"  after {steps} steps:")
        ok &= cmp("density", gpu.density(), orc.density())
        ok &= cmp("velocity", gpu.getVelocity(), orc.velocities())
        ok &= cmp("position", gpu.getParticles(), orc.positions())
    gpu.close()
    orc.close()
    print("  RESULT", "BIT-IDENTICAL" if ok else "DIFFERS")
    return ok


def timing(scene, res, steps=20):
    p = sf.default_params(res, scene)
    pos = sf.scene_generate(p)
    gpu = sf.SPHSolver(p)
    gpu.setParticles(pos)
    gpu.makeReady()
    gpu.advanceSteps(5)
    gpu.synchronize()
    gpu.profileEnable(True)
    gpu.profileReset()
    gpu.timerStart()
    gpu.advanceSteps(steps)
    ms = gpu.timerStop()
    prof = gpu.profile()
    n = len(pos)
    print(f"== timing {scene} res {res} N={n}: {ms / steps:.3f} ms/step  {n * steps / ms * 1e3:.3e} particle-steps/s")
    for k, (t, c) in prof.items():
        if c:
            print(f"   {k:18s} {t / steps:8.3f} ms/step  ({c // steps} launches/step)")
    gpu.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        check("Dambreak", 24, 2)
        sys.exit(0)
    allok = True
    allok &= check("Dambreak", 24, 30)
    allok &= check("CubeDrop", 24, 5)
    allok &= check("SphereDrop", 24, 5)
    allok &= check("DoubleDambreak", 24, 5)
    allok &= check("Dambreak", 24, 3, bUseAttractivePressure=1)
    allok &= check("Dambreak", 24, 3, bCorrectDensity=1)
    allok &= check("Dambreak", 61, 3)
    print("ALL", "OK" if allok else "FAILED")
    timing("CubeDrop", 100)
    timing("Dambreak", 203)
