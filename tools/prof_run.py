"""Tiny driver for ncu captures: python tools/prof_run.py <scene> <res> <steps> [warm]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplefluid_b200 as sf  # noqa: E402

scene, res, steps = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 0
p = sf.default_params(res, scene)
pos = sf.scene_generate(p)
g = sf.SPHSolver(p)
g.setParticles(pos)
g.makeReady()
if warm:
    g.advanceSteps(warm)
g.synchronize()
g.timerStart()
g.advanceSteps(steps)
ms = g.timerStop()
print(f"{scene} res {res} N={len(pos)} {ms/steps:.3f} ms/step")
g.close()
