"""Generates tests/golden/exe_*.npz: outputs of the REFERENCE'S OWN compiled SPH step.

    python tests/golden/make_exe_golden.py        (build container only: needs /root/reference/Prebuild/SimpleFluid.exe)

oracle/exe/sf_exe_harness.c maps the reference's shipped binary and calls its SPHSolver::makeReady (EXE@0x140016650) and
SPHSolver::advanceFrame (EXE@0x140016810) natively (header of that file: how).  /root/reference does not exist on the
GPU box, so these fixtures are how the reference's real outputs travel: the oracle (tests/test_oracle.py) and the CUDA
path (tests/test_parity_gpu.py) are both compared with them bit for bit.  Inputs are the scenes of the reference's
Source/SceneManager.cpp (pinned separately by scene_checksums.json); the wall-particle jitter seed -- std::random_device
in the reference -- is the `seed` recorded in each file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import exe_harness as eh  # noqa: E402
import oracle_binding as ob  # noqa: E402

# name: (scene, resolution, substeps, seed, parameter overrides, substeps whose full fields are stored)
CASES = {
    "dambreak24": ("Dambreak", 24, 1000, 0, {}, (0, 1, 2, 999)),            # reference default; the CFL branch of dt is taken
    "spheredrop24": ("SphereDrop", 24, 400, 7, {}, (0, 399)),
    "doubledambreak24": ("DoubleDambreak", 24, 200, 3, {}, (0, 199)),
    "cubedrop16_shepard": ("CubeDrop", 16, 300, 0, {"bCorrectDensity": 1}, (0, 1, 299)),
    "dambreak16_attractive_nowalls": ("Dambreak", 16, 300, 0, {"bUseAttractivePressure": 1, "bUseBoundaryParticles": 0}, (0, 299)),
    "cubedrop61_oddgrid": ("CubeDrop", 61, 2, 11, {}, (0, 1)),               # ceilf(2/h) = 62 cells per axis
}


def main():
    assert eh.available(), "needs /root/reference/Prebuild/SimpleFluid.exe"
    for name, (scene, res, steps, seed, over, keep) in CASES.items():
        p = ob.default_params(res, scene)
        for k, v in over.items():
            setattr(p, k, v)
        pos = ob.scene(p)
        E = eh.run(p, pos, steps, seed=seed)
        assert all(o == 1 for o in E["ordered"]) and all(m == len(pos) for m in E["listed"])
        out = {"scene": scene, "resolution": res, "seed": seed, "overrides": repr(over), "n": len(pos), "grid": np.array(E["grid"]),
               "dts": np.array(E["dt"], np.float32), "steps_kept": np.array(keep),
               "cubic_consts": E["cubic_consts"], "spiky_consts": E["spiky_consts"], "particleMass": np.float32(E["particleMass"])}
        if name == "dambreak24":
            out["cubic_W"], out["spiky_gradW"] = E["cubic_W"], E["spiky_gradW"]
            for w in range(6):
                out[f"wall{w}"] = E["walls"][w]
        for k in keep:
            for f in ("cell", "rho", "acc", "x", "v"):
                out[f"{f}{k}"] = E[f][k]
        np.savez_compressed(os.path.join(HERE, f"exe_{name}.npz"), **out)
        print(f"exe_{name}.npz: N={len(pos)} steps={steps} dt in [{min(E['dt']):.6g}, {max(E['dt']):.6g}]")


if __name__ == "__main__":
    main()
