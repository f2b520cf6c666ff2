"""Full-size checksums of the REFERENCE BINARY's own outputs: tests/golden/exe_fullsize_checksums.json.

    python tests/golden/make_exe_fullsize.py [case ...]       (build container only; C5 needs ~15 GB of RAM and ~10 minutes)

For every BASELINE.json configuration at its full size -- C2 CubeDrop 1 M, C3 DoubleDambreak 8.03 M, C4 SphereDrop
16.05 M, the weak-scaling unit Dambreak 8.09 M and C5 Dambreak 64.2 M -- the reference's own compiled makeReady /
advanceFrame (oracle/exe/sf_exe_harness.c over Prebuild/SimpleFluid.exe) advances the scene by one or two substeps; the
SHA-256 of every output field (cell index, density, acceleration, position, velocity, as raw little-endian arrays in
particle order) and dt are recorded.  The fields themselves would be gigabytes; the checksums travel.  The GPU tests hash
what the CUDA path produces for the same scene and compare (tests/test_parity_gpu.py)."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import exe_harness as eh  # noqa: E402
import oracle_binding as ob  # noqa: E402

CASES = {  # name: (scene, resolution, substeps)
    "C2_cubedrop_1m": ("CubeDrop", 100, 2),
    "C3_doubledambreak_8m": ("DoubleDambreak", 161, 1),
    "weak_unit_dambreak_8m": ("Dambreak", 203, 2),
    "C4_spheredrop_16m": ("SphereDrop", 313, 1),
    "C5_dambreak_64m": ("Dambreak", 404, 1),
}
OUT = os.path.join(HERE, "exe_fullsize_checksums.json")


def digest(a):
    return hashlib.sha256(memoryview(a).cast("B")).hexdigest()


def verify_oracle():
    """Second pass (no binary needed): the CPU oracle on every recorded case; stores `oracle_bit_identical`."""
    out = json.load(open(OUT))
    for name, rec in out["cases"].items():
        p = ob.default_params(rec["resolution"], rec["scene"])
        pos = ob.scene(p)
        orc = ob.Oracle(p, pos, boundary_seed=0)
        ok = True
        for k, want in enumerate(rec["steps"]):
            dt = orc.advance()
            got = {"cell": orc.cell_index(), "rho": orc.density(), "acc": orc.accel(), "x": orc.positions(), "v": orc.velocities()}
            ok = ok and float(dt) == rec["dts"][k] and all(digest(got[f]) == want[f] for f in want)
        orc.close()
        rec["oracle_bit_identical"] = bool(ok)
        json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
        print(f"{name}: oracle bit-identical to the binary: {ok}", flush=True)


def main():
    if sys.argv[1:] == ["--verify-oracle"]:
        return verify_oracle()
    assert eh.available(), "needs /root/reference/Prebuild/SimpleFluid.exe"
    out = json.load(open(OUT)) if os.path.exists(OUT) else {"source": "Prebuild/SimpleFluid.exe executed by oracle/exe/sf_exe_harness.c", "seed": 0, "cases": {}}
    for name in (sys.argv[1:] or CASES):
        scene, res, steps = CASES[name]
        p = ob.default_params(res, scene)
        pos = ob.scene(p)
        t0 = time.time()
        E = eh.run(p, pos, steps, seed=0)
        rec = {"scene": scene, "resolution": res, "n": len(pos), "grid": list(E["grid"]), "dts": [float(d) for d in E["dt"]], "steps": []}
        for k in range(steps):
            assert E["ordered"][k] == 1 and E["listed"][k] == len(pos)
            rec["steps"].append({f: digest(E[f][k]) for f in ("cell", "rho", "acc", "x", "v")})
        out["cases"][name] = rec
        json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
        print(f"{name}: N={len(pos)} {steps} substeps in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
