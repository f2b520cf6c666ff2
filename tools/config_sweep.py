"""Throughput of every BASELINE.json config on one GPU (device-resident, CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplefluid_b200 as sf  # noqa: E402

CONFIGS = [("C1 dambreak default", "Dambreak", 24, 400), ("C2 cube 1M", "CubeDrop", 100, 100),
           ("C3 doubledambreak 8M", "DoubleDambreak", 161, 40), ("C4 sphere 16M", "SphereDrop", 313, 30),
           ("8M dambreak (bench unit)", "Dambreak", 203, 40)]
if len(sys.argv) > 1 and sys.argv[1] == "c5":  # the whole 64M-particle weak-scaling scene on ONE GPU (maximum size)
    CONFIGS = [("C5 dambreak 64M on 1 GPU", "Dambreak", 404, 8)]
for name, scene, res, steps in CONFIGS:
    p = sf.default_params(res, scene)
    pos = sf.scene_generate(p)
    g = sf.SPHSolver(p)
    g.setParticles(pos)
    g.makeReady()
    g.advanceSteps(3 if len(pos) > 30_000_000 else 10)
    g.synchronize()
    g.timerStart()
    g.advanceSteps(steps)
    ms = g.timerStop()
    print(f"{name:28s} N={len(pos):9d}  {ms / steps:8.4f} ms/substep  {len(pos) * steps / ms * 1e3:.3e} particle-steps/s", flush=True)
    g.close()
