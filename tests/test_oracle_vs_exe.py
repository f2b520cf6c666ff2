"""Pins the oracle to the reference MECHANICALLY.

1. Committed fixtures tests/golden/exe_*.npz hold outputs of the reference's OWN compiled SPH step (Prebuild/SimpleFluid.exe
   executed natively by oracle/exe/sf_exe_harness.c -- see its header and tests/golden/make_exe_golden.py).  The oracle must
   reproduce them bit for bit: dt sequence, cell indices, density, acceleration, positions, velocities, kernel tables, wall
   particles -- for every scene, the parameter variants, an odd grid, and 1000 substeps of the reference default.  These run
   everywhere (also on the GPU box, where /root/reference does not exist).
2. Where the reference is present (the build container), the binary is executed LIVE on further inputs (random states with
   non-zero velocities, other seeds and resolutions), and static facts the oracle relies on are read from the image: the
   .rdata constants at the addresses SURVEY.md Appendix D cites, the import-table slots behind the math thunks, and the
   role of each random draw in generateBoundaryParticles (EXE@0x14001705a-0x1400172e0)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import exe_golden
import exe_harness as eh

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
needs_exe = pytest.mark.skipif(not eh.available(), reason="the reference binary is only present in the build container")


def oracle_for(ob, g):
    p = ob.default_params(g["resolution"], g["scene"], **g["overrides"])
    pos = ob.scene(p)
    assert len(pos) == int(g["n"])
    return p, pos, ob.Oracle(p, pos, boundary_seed=g["seed"])


@pytest.mark.parametrize("name", exe_golden.CASES)
def test_oracle_reproduces_the_reference_binarys_outputs(ob, name):
    g = exe_golden.load(name)
    p, pos, orc = oracle_for(ob, g)
    assert tuple(orc.grid_dims()) == tuple(int(x) for x in g["grid"])
    assert np.float32(p.particleMass) == g["particleMass"]
    kc = orc.kernel_consts()  # W_zero, radius2, invStep, ... of the cubic kernel
    assert {float(g["cubic_consts"][3]), float(g["cubic_consts"][1]), float(g["cubic_consts"][2])} <= set(float(x) for x in kc)
    if "cubic_W" in g:
        assert np.array_equal(orc.table(0)[:10000], g["cubic_W"]), "cubic W table differs from PrecomputedKernel<Cubic> of the binary"
        assert np.array_equal(orc.table(1), g["spiky_gradW"]), "spiky gradW table differs from the binary's"
        for w in range(6):
            assert np.array_equal(orc.boundary(w), g[f"wall{w}"]), f"wall {w} particles differ from generateBoundaryParticles of the binary"
    n = exe_golden.replay(g, orc.advance, lambda: dict(cell=orc.cell_index(), rho=orc.density(), acc=orc.accel(), x=orc.positions(), v=orc.velocities()))
    assert n == len(g["dts"])
    if name == "dambreak24":
        assert g["dts"].min() < np.float32(1e-3), "the CFL branch of computeTimeStep must be exercised"
    orc.close()


def test_oracle_matches_the_binary_at_full_size_c2(ob):
    """BASELINE.json configs[1] at full size (CubeDrop, 1,000,000 particles), two substeps: SHA-256 of every field against
    the checksums of the reference binary's own outputs (tests/golden/make_exe_fullsize.py).  The larger configurations
    (8 M ... 64 M particles) are checked the same way on the GPU (tests/test_parity_gpu.py)."""
    import hashlib
    import json
    rec = json.load(open(os.path.join(HERE, "golden", "exe_fullsize_checksums.json")))["cases"]["C2_cubedrop_1m"]
    p = ob.default_params(rec["resolution"], rec["scene"])
    pos = ob.scene(p)
    assert len(pos) == rec["n"] == 1000000
    orc = ob.Oracle(p, pos, boundary_seed=0)
    digest = lambda a: hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()
    for k, want in enumerate(rec["steps"]):
        assert np.float32(orc.advance()) == np.float32(rec["dts"][k])
        got = {"cell": orc.cell_index(), "rho": orc.density(), "acc": orc.accel(), "x": orc.positions(), "v": orc.velocities()}
        for f in want:
            assert digest(got[f]) == want[f], f"{f} after substep {k}"
    orc.close()


# ---- live runs of the binary (build container only) ---------------------------------------------------------------
def compare_live(ob, p, pos, vel, steps, seed):
    E = eh.run(p, pos, steps, seed=seed, vel=vel)
    orc = ob.Oracle(p, pos, vel, boundary_seed=seed)
    assert all(o == 1 for o in E["ordered"]), "the binary's cell lists are not in ascending particle order"
    for k in range(steps):
        assert orc.advance() == E["dt"][k], f"dt of substep {k}"
        for name, a in (("cell", orc.cell_index()), ("rho", orc.density()), ("acc", orc.accel()), ("x", orc.positions()), ("v", orc.velocities())):
            assert a.tobytes() == np.asarray(E[name][k]).tobytes(), f"{name} after substep {k} differs from the reference binary"
    orc.close()


@needs_exe
@pytest.mark.parametrize("scene,res,steps,seed,over", [
    ("Dambreak", 12, 120, 1, {}), ("SphereDrop", 20, 60, 2, {}), ("CubeDrop", 13, 40, 3, {}), ("DoubleDambreak", 16, 40, 4, {}),
    ("Dambreak", 12, 60, 5, {"bCorrectDensity": 1, "bUseAttractivePressure": 1}), ("Dambreak", 14, 40, 6, {"bUseBoundaryParticles": 0}),
    ("Dambreak", 12, 40, 0, {"pressureStiffness": 20000.0, "viscosity": 0.2, "boundaryRestitution": 0.5}),
])
def test_live_binary_vs_oracle_scenes(ob, scene, res, steps, seed, over):
    p = ob.default_params(res, scene, **over)
    compare_live(ob, p, ob.scene(p), None, steps, seed)


@needs_exe
def test_live_binary_vs_oracle_random_moving_state(ob):
    """Ragged cells, particles against every wall, fast particles (CFL-limited dt, wall bounces with restitution)."""
    rng = np.random.default_rng(5)
    p = ob.default_params(10, "Dambreak")
    r = p.particleRadius
    pos = (rng.random((3000, 3)) * (2 - 2 * r) - (1 - r)).astype(np.float32)
    pos[:400] = np.float32(1 - r) * np.sign(pos[:400])  # corners and walls, clamped positions
    vel = (rng.standard_normal((3000, 3)) * 3).astype(np.float32)
    compare_live(ob, p, pos, vel, 60, 9)


# ---- static facts read from the image ------------------------------------------------------------------------------
class Image:
    def __init__(self):
        self.data = open(eh.EXE, "rb").read()
        pe = struct.unpack_from("<I", self.data, 0x3c)[0]
        nsec, optsz = struct.unpack_from("<H", self.data, pe + 6)[0], struct.unpack_from("<H", self.data, pe + 20)[0]
        self.base = struct.unpack_from("<Q", self.data, pe + 24 + 24)[0]
        self.secs = []
        for i in range(nsec):
            o = pe + 24 + optsz + 40 * i
            vsize, va, rawsize, rawptr = struct.unpack_from("<IIII", self.data, o + 8)
            self.secs.append((va, max(vsize, rawsize), rawptr))

    def at(self, va, fmt):
        rva = va - self.base
        for v, size, raw in self.secs:
            if v <= rva < v + size:
                return struct.unpack_from(fmt, self.data, raw + rva - v)[0]
        raise KeyError(hex(va))

    def disasm(self, start, stop):
        out = subprocess.run(["objdump", "-d", "-M", "intel", "--no-show-raw-insn", f"--start-address={start:#x}", f"--stop-address={stop:#x}", eh.EXE],
                             capture_output=True, text=True, check=True).stdout
        return [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"^\s*([0-9a-f]+):\s+(.*)$", out, re.M)]


@pytest.fixture(scope="module")
def image():
    return Image()


@needs_exe
def test_rdata_constants_the_oracle_hard_codes(image):
    """SURVEY.md Appendix D addresses; each value appears literally in oracle/sf_oracle.c."""
    src = open(os.path.join(ROOT, "oracle", "sf_oracle.c")).read()
    f32 = lambda v: struct.unpack("<f", struct.pack("<f", v))[0]
    consts = [  # (VA, format, value, literal in the oracle source)
        (0x1400561d0, "<d", 1e-8, "1e-8"), (0x1400561d8, "<f", f32(0.2), "0.2f"), (0x1400561c4, "<f", f32(0.1), "0.1f"),
        (0x140056308, "<f", 10.0, "10.0f"), (0x140056310, "<d", 1e10, "1e10"), (0x140056290, "<d", 9.8, "9.8"),
        (0x140056258, "<d", 1.0, "1.0"), (0x140056278, "<d", 6.0, "6.0"), (0x140056270, "<d", 3.0, "3.0"), (0x140056288, "<d", 8.0, "8.0"),
        (0x1400562c0, "<d", 48.0, "48.0"), (0x1400562a8, "<d", 15.0, "15.0"), (0x140056378, "<d", -45.0, "45.0"),
        (0x1400562c8, "<f", f32(3.14159274), "3.14159274f"), (0x140056334, "<f", 10000.0, "10000.0f"), (0x140056228, "<d", 0.1, "0.1"),
        (0x140056298, "<d", 10.0, "10.0"), (0x140056238, "<d", 0.3, "0.3"), (0x140056240, "<f", f32(1.7), "1.7f"), (0x1400562b8, "<f", 3.0, "3.0f"),
        (0x1400561f8, "<f", 0.5, "0.5f"), (0x14005620c, "<f", 1.0, "1.0f"), (0x140056250, "<d", 0.9, "0.9"), (0x140056338, "<f", 4294967296.0, "4294967296.0f"),
    ]
    for va, fmt, value, literal in consts:
        assert image.at(va, fmt) == value, f"constant at {va:#x}"
        assert literal in src, f"the oracle does not spell {literal}"


@needs_exe
def test_math_thunks_bind_to_the_expected_imports(image):
    """The thunks the step calls (EXE@0x14003fc24 ceilf, 0x14003fc30 floorf, 0x14003fc48 sqrtf ...) jump through these IAT slots."""
    out = subprocess.run(["objdump", "-p", eh.EXE], capture_output=True, text=True, check=True).stdout
    block = out[out.index("DLL Name: api-ms-win-crt-math-l1-1-0.dll"):]
    names = re.findall(r"^\t[0-9a-f]+\t\s+\d+\s+(\w+)$", block.split("\n\n")[0], re.M)
    first_thunk = int(re.findall(r"([0-9a-f]{8})\n\n\tDLL Name: api-ms-win-crt-math", out)[0], 16)
    slots = {name: image.base + first_thunk + 8 * i for i, name in enumerate(names)}
    assert slots["fminf"] == 0x140048088 and slots["fmaxf"] == 0x140048090 and slots["ceilf"] == 0x140048098 and slots["fmax"] == 0x1400480a0
    assert slots["powf"] == 0x1400480b0 and slots["pow"] == 0x1400480c8 and slots["floorf"] == 0x1400480d0 and slots["sqrtf"] == 0x1400480e8
    for thunk, name in ((0x14003fc24, "ceilf"), (0x14003fc30, "floorf"), (0x14003fc36, "pow"), (0x14003fc3c, "powf"), (0x14003fc48, "sqrtf")):
        (_, ins), = image.disasm(thunk, thunk + 6)
        assert ins.startswith("jmp") and ins.endswith(f"# {slots[name]:#x}"), (hex(thunk), ins)


@needs_exe
def test_wall_particle_draw_roles(image):
    """generateBoundaryParticles, EXE@0x14001705a-0x1400172e0: per wall three generate_canonical draws (call 0x1400101a0);
    a draw scaled by xmm9 (= hi - lo) and offset by xmm12 (= lo) is a tangential jitter, one scaled by xmm10 (= lo - 0) and
    offset by xmm13 (= 0) the depth jitter.  Sequence of roles per wall, and which lattice coordinate (xmm15 = ti of the
    outer loop, xmm11 = tj of the middle loop) or box face ([rax+0x10..0x24] = boxMin/boxMax) each lands on:
    oracle/sf_oracle.c sfo_generate_boundary and csrc/sf_host.cpp generate_boundary follow this table."""
    ins = image.disasm(0x14001705a, 0x1400172e0)
    walls, cur = [], None
    for _, s in ins:
        if s.startswith("call") and "0x1400101a0" in s:
            if cur is None or len(cur) == 3 and cur[-1] is not None:
                cur = []
                walls.append(cur)
            cur.append(None)
        m = re.match(r"mulss\s+xmm\d+,xmm(9|10)$", s)
        if m and cur is not None and cur and cur[-1] is None:
            cur[-1] = "J" if m.group(1) == "9" else "D"
    assert walls == [["J", "J", "D"], ["J", "J", "D"], ["J", "D", "J"], ["J", "D", "J"], ["D", "J", "J"], ["D", "J", "J"]]
    text = "\n".join(s for _, s in ins)
    faces = re.findall(r"(?:movss\s+xmm1,|addss\s+xmm[18],)DWORD PTR \[rax\+(0x[0-9a-f]+)\]", text)
    assert faces == ["0x10", "0x1c", "0x14", "0x20", "0x18", "0x24"]  # boxMin.x, boxMax.x, boxMin.y, boxMax.y, boxMin.z, boxMax.z
    # generate_canonical<float,24> (EXE@0x1400101a0): one draw, converted by cvtsi2ss of the zero-extended 32-bit value, divided by 2^32
    gc = "\n".join(s for _, s in image.disasm(0x1400101a0, 0x1400102e2))
    assert "cvtsi2ss xmm1,rax" in gc and gc.count("divss") == 2 and "0x140056338" in gc
