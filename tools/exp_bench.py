"""Kernel experiments: per-kernel times of one library build on the developed dambreak (development aid).
    SF_B200_LIB=/path/to/variant.so python tools/exp_bench.py [res] [settle] [steps]
Prints per-kernel ms and a checksum of the final state (every variant must print the same one)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplefluid_b200 as sf  # noqa: E402

res = float(sys.argv[1]) if len(sys.argv) > 1 else 203
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
scene = sys.argv[4] if len(sys.argv) > 4 else "Dambreak"
p = sf.default_params(res, scene)
pos = sf.scene_generate(p)
g = sf.SPHSolver(p)
g.setParticles(pos)
g.generateBoundaryParticles(0)
g.makeReady()
g.advanceSteps(10)
g.synchronize()
g.timerStart(); g.advanceSteps(steps); ms_rest = g.timerStop()
g.advanceSteps(settle)
g.synchronize()
g.timerStart(); g.advanceSteps(steps); ms = g.timerStop()
g.profileEnable(True, every=1)
g.profileReset()
g.advanceSteps(16)
prof = g.profile()
g.profileEnable(False)
d = g.diagnostics()
h = hashlib.sha256(g.getParticles().tobytes() + g.getVelocity().tobytes()).hexdigest()[:16]
print(f"{os.path.basename(sf.library_path())}: {scene} res {res} N={len(pos)} at rest {ms_rest / steps:.3f} ms/step, developed {ms / steps:.3f} ms/step "
      f"= {len(pos) * steps / ms * 1e3:.3e} p-steps/s; fallback bricks {d['fallback_bricks']} nbr_mean {d['nbr_mean']:.1f} state {h}")
print("   " + "  ".join(f"{k[2:]}={t / c:.3f}" for k, (t, c) in prof.items() if c))
dbg = g.debugCounters()
if dbg[3]:
    print(f"   density consumer warps: waiting for a brick {dbg[0] / dbg[3] * 100:.1f} %, exact phase {dbg[2] / dbg[3] * 100:.1f} % of their cycles")
    if dbg[1]:
        print(f"   producers (all pair kernels), per refill: {dbg[4] / dbg[1]:.0f} cycles waiting for a free staging buffer; {dbg[1]} refills")
g.close()
