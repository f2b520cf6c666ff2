"""Compare the selectable kernel variants (SF_SORT=radix|count) on one GPU: bit-equality of
the state after `steps` substeps against the first variant, per-kernel times, and the host-buffer step.
Development aid; usage:  python tools/variant_bench.py [res] [steps] [variant ...]  where a variant is
the sort, count (the default) or radix."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplefluid_b200 as sf  # noqa: E402


def make(scene, res, variant):
    os.environ["SF_SORT"] = variant.split("+")[0]
    p = sf.default_params(res, scene)
    pos = sf.scene_generate(p)
    gpu = sf.SPHSolver(p)  # the variant is read at sf_create
    gpu.setParticles(pos)
    gpu.generateBoundaryParticles(0)
    gpu.makeReady()
    return gpu, len(pos)


def run(scene, res, steps, variants):
    ref = None
    for v in variants:
        gpu, n = make(scene, res, v)
        gpu.advanceSteps(steps)
        gpu.synchronize()
        state = (gpu.getParticles().copy(), gpu.getVelocity().copy(), gpu.density().copy())
        same = "reference" if ref is None else ("BIT-IDENTICAL" if all(np.array_equal(a, b) for a, b in zip(state, ref)) else "DIFFERS")
        if ref is None:
            ref = state
        gpu.profileEnable(True)
        gpu.profileReset()
        gpu.timerStart()
        gpu.advanceSteps(steps)
        ms = gpu.timerStop()
        prof = gpu.profile()
        gpu.profileEnable(False)
        gpu.timerStart()
        gpu.advanceSteps(steps)  # CUDA-graph replay, no per-kernel events
        ms_graph = gpu.timerStop()
        print(f"== {scene} res {res} N={n} variant {v}: {same}; {ms / steps:.3f} ms/step profiled, {ms_graph / steps:.3f} ms/step graph "
              f"= {n * steps / ms_graph * 1e3:.3e} particle-steps/s", flush=True)
        print("   " + "  ".join(f"{k[2:]}={t / steps:.3f}" for k, (t, c) in prof.items() if c), flush=True)
        gpu.close()


def e2e(scene, res, steps, variant):
    gpu, n = make(scene, res, variant)
    gpu.advanceSteps(3)
    hx = sf.PinnedArray((n, 3))
    hv = sf.PinnedArray((n, 3))
    hx.array[:] = gpu.getParticles()
    hv.array[:] = gpu.getVelocity()
    # equality of the host-buffer step with the resident path
    x0, v0 = hx.array.copy(), hv.array.copy()
    gpu.stepHost(hx.array, hv.array)
    gpu.stepHost(hx.array, hv.array)
    chk, _ = make(scene, res, variant)
    chk.setParticles(x0, v0)
    chk.makeReady()
    chk.advanceSteps(2)
    ok = np.array_equal(chk.getParticles(), hx.array) and np.array_equal(chk.getVelocity(), hv.array)
    chk.close()
    t0 = time.perf_counter()
    for _ in range(steps):
        gpu.stepHost(hx.array, hv.array)
    dt = time.perf_counter() - t0
    print(f"== e2e {scene} res {res} N={n} variant {variant}: host-buffer step {'BIT-IDENTICAL' if ok else 'DIFFERS'} to resident; "
          f"{dt / steps * 1e3:.3f} ms/step = {n * steps / dt:.3e} particle-steps/s ({48 * n / 1e6:.0f} MB moved per step)", flush=True)
    gpu.close()
    hx.close()
    hv.close()


if __name__ == "__main__":
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 203
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    variants = sys.argv[3:] or ["count", "radix"]
    run("Dambreak", res, steps, variants)
    e2e("Dambreak", res, 10, variants[0])
