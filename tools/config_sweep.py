"""Throughput of every BASELINE.json config on one GPU (device-resident, CUDA events, graph replay): at rest (the initial
lattice) and on the developed flow after `settle` substeps.
    python tools/config_sweep.py [all|c5] [settle]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplefluid_b200 as sf  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
CONFIGS = [("C1 dambreak default", "Dambreak", 24, 400), ("C2 cube 1M", "CubeDrop", 100, 100),
           ("C3 doubledambreak 8M", "DoubleDambreak", 161, 40), ("C4 sphere 16M", "SphereDrop", 313, 30),
           ("8M dambreak (bench unit)", "Dambreak", 203, 40)]
if which == "c5":  # the whole 64M-particle weak-scaling scene on ONE GPU (maximum size)
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
    line = f"{name:28s} N={len(pos):9d}  at rest {ms / steps:8.4f} ms/substep {len(pos) * steps / ms * 1e3:.3e} p-steps/s"
    if settle:
        g.advanceSteps(settle)
        g.synchronize()
        g.timerStart()
        g.advanceSteps(steps)
        ms = g.timerStop()
        d = g.diagnostics()
        line += (f"  | after {settle} substeps {ms / steps:8.4f} ms/substep {len(pos) * steps / ms * 1e3:.3e} p-steps/s, {d['nbr_mean']:.1f} neighbours/particle"
                 f" (max {d['nbr_max']}), fallback bricks {d['fallback_bricks']}, particles without list {d['particles_without_list']}")
    print(line, flush=True)
    g.close()
