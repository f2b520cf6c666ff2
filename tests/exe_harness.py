"""Test infrastructure: drives oracle/_ref/sf_exe_harness, which runs the reference's OWN compiled SPH step
(Prebuild/SimpleFluid.exe: makeReady EXE@0x140016650, advanceFrame EXE@0x140016810) natively, and parses its dump.
Only available where /root/reference exists (the build container); the GPU box sees its outputs as the committed
fixtures tests/golden/exe_*.npz (made by tests/golden/make_exe_golden.py)."""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = "/root/reference/Prebuild/SimpleFluid.exe"
HARNESS = os.path.join(ROOT, "oracle", "_ref", "sf_exe_harness")
SOURCE = os.path.join(ROOT, "oracle", "exe", "sf_exe_harness.c")

FLAG_CORRECT_DENSITY, FLAG_BOUNDARY, FLAG_ATTRACTIVE, FLAG_VELOCITIES = 1, 2, 4, 8


def available():
    return os.path.exists(EXE) and os.path.exists(SOURCE)


def build():
    if not os.path.exists(HARNESS) or os.path.getmtime(HARNESS) < os.path.getmtime(SOURCE):
        os.makedirs(os.path.dirname(HARNESS), exist_ok=True)
        subprocess.run(["gcc", "-O1", "-Wall", "-o", HARNESS, SOURCE, "-lm"], check=True)
    return HARNESS


def run(params, pos, steps, seed=0, vel=None):
    """params: any object with the sf_params field names (the product's SFParams or the oracle's).  Returns a dict of
    arrays: per-step lists `dt`, `cell`, `rho`, `acc`, `x`, `v`, plus `walls`, the kernel tables and the grid."""
    build()
    pos = np.ascontiguousarray(pos, np.float32)
    n = len(pos)
    flags = (FLAG_CORRECT_DENSITY if params.bCorrectDensity else 0) | (FLAG_BOUNDARY if params.bUseBoundaryParticles else 0) | \
            (FLAG_ATTRACTIVE if params.bUseAttractivePressure else 0) | (FLAG_VELOCITIES if vel is not None else 0)
    with tempfile.TemporaryDirectory(prefix="sfexe") as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<4I", n, steps, seed, flags))
            f.write(struct.pack("<8f", params.kernelRadius, params.pressureStiffness, params.viscosity, params.boundaryRestitution,
                                params.attractivePressureRatio, params.restDensity, params.defaultTimestep, 0.0))
            f.write(pos.tobytes())
            if vel is not None:
                f.write(np.ascontiguousarray(vel, np.float32).tobytes())
        r = subprocess.run([HARNESS, EXE, fin, fout], capture_output=True, text=True, timeout=3600)
        if r.returncode:
            raise RuntimeError(f"sf_exe_harness failed ({r.returncode}): {r.stderr[-2000:]}")
        raw = open(fout, "rb").read()
    return parse(raw)


def parse(raw):
    o = 0

    def take(dtype, count):
        nonlocal o
        a = np.frombuffer(raw, dtype, count, o)
        o += a.nbytes
        return a

    hdr = take(np.uint32, 16)
    assert hdr[0] == 0x45584553
    n, steps = int(hdr[1]), int(hdr[2])
    out = {"n": n, "steps": steps, "grid": tuple(int(x) for x in hdr[3:6]), "wall_counts": [int(x) for x in hdr[6:12]]}
    out["params_raw"] = bytes(take(np.uint8, 0x5c))
    pf = np.frombuffer(out["params_raw"], np.float32)
    out["particleMass"], out["particleRadius"], out["kernelRadiusSqr"] = float(pf[0x48 // 4]), float(pf[0x4c // 4]), float(pf[0x50 // 4])
    out["cubic_hklW0"] = take(np.float32, 4)
    out["cubic_consts"] = take(np.float32, 4)   # radius, radius2, invStep, W_zero
    out["cubic_W"] = take(np.float32, 10000)
    out["spiky_hklW0"] = take(np.float32, 4)
    out["spiky_consts"] = take(np.float32, 4)
    out["spiky_gradW"] = take(np.float32, 10001)
    out["walls"] = [take(np.float32, 3 * c).reshape(-1, 3) for c in out["wall_counts"]]
    for k in ("dt", "cell", "rho", "acc", "x", "v", "listed", "ordered", "ncells"):
        out[k] = []
    for _ in range(steps):
        out["dt"].append(float(take(np.float32, 1)[0]))
        rec = take(np.uint32, 3)
        out["listed"].append(int(rec[0]))
        out["ordered"].append(int(rec[1]))
        out["ncells"].append(int(rec[2]))
        out["cell"].append(take(np.uint32, n))
        out["rho"].append(take(np.float32, n))
        out["acc"].append(take(np.float32, 3 * n).reshape(n, 3))
        out["x"].append(take(np.float32, 3 * n).reshape(n, 3))
        out["v"].append(take(np.float32, 3 * n).reshape(n, 3))
    assert o == len(raw)
    return out
