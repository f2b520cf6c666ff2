"""CPU baseline of BASELINE.md section 3: the oracle port timed on this box's host cores (all threads and 1 thread)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402

print(f"host cores: {os.cpu_count()}")
print("| config | N | threads | substeps | s/substep | particle-steps/s | collect (serial) share |")
print("|---|---|---|---|---|---|---|")
for name, scene, res, steps in (("C1 Dambreak res 24", "Dambreak", 24, 200), ("C2 CubeDrop res 100", "CubeDrop", 100, 20),
                                ("C3 DoubleDambreak res 161", "DoubleDambreak", 161, 3)):
    p = ob.default_params(res, scene)
    pos = ob.scene(p)
    for threads in (0, 1):
        if threads == 1 and len(pos) > 2_000_000:
            continue
        orc = ob.Oracle(p, pos, boundary_seed=0, threads=threads)
        orc.advance()
        t0 = time.perf_counter()
        collect = 0.0
        for _ in range(steps):
            orc.advance()
            collect += orc.timing()[1]
        dt = time.perf_counter() - t0
        print(f"| {name} | {len(pos)} | {threads or os.cpu_count()} | {steps} | {dt / steps:.4f} | {len(pos) * steps / dt:.3e} | {collect / dt * 100:.1f}% |", flush=True)
        orc.close()
