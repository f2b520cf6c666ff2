"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol declared
in include/sf_b200.h, and its host-side setup (parameters, scenes, kernel tables, wall particles)
agrees bit for bit with the golden vectors and the oracle.  No compute calls without a GPU."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(sf):
    lib = C.CDLL(sf.library_path())
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    from simplefluid_b200 import binding
    assert set(binding.EXPORTS) == set(names)


def test_params_defaults_and_update(sf, ob):
    p, po = sf.default_params(24, "Dambreak"), ob.default_params(24, "Dambreak")
    for name, _ in sf.SFParams._fields_:
        a, b = getattr(p, name), getattr(po, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name
    p.kernelRadius = 2.0 / 64.0
    p.updateParams()
    po = ob.default_params(64, "Dambreak")
    assert p.particleMass == po.particleMass and p.particleRadius == po.particleRadius and p.kernelRadiusSqr == po.kernelRadiusSqr


@pytest.mark.parametrize("res", [24, 48, 100])
@pytest.mark.parametrize("scene", ["SphereDrop", "CubeDrop", "Dambreak", "DoubleDambreak"])
def test_scene_generate_matches_reference_golden(sf, scene, res):
    with open(os.path.join(HERE, "golden", "scene_checksums.json")) as f:
        g = json.load(f)["scenes"][f"{scene}@{res}"]
    pos = sf.scene_generate(sf.default_params(res, scene))
    assert len(pos) == g["n"]
    assert hashlib.sha256(pos.tobytes()).hexdigest() == g["sha256"]


def test_scene_generate_count_query_and_truncation(sf):
    p = sf.default_params(24, "CubeDrop")
    n = C.c_uint64(0)
    assert sf.library().sf_scene_generate(C.byref(p), 1, None, 0, C.byref(n)) == 0 and n.value == 13824
    buf = np.full((10, 3), 7.0, np.float32)
    assert sf.library().sf_scene_generate(C.byref(p), 1, buf.ctypes.data, 5, C.byref(n)) == 0 and n.value == 13824
    assert np.all(buf[5:] == 7.0) and np.all(buf[:5] != 7.0)
    assert sf.library().sf_scene_generate(C.byref(p), 9, None, 0, C.byref(n)) != 0


@pytest.mark.parametrize("res", [24, 61, 100, 203])
def test_kernel_tables_match_oracle(sf, ob, res):
    from simplefluid_b200 import binding
    w, g, c = binding.build_tables(sf.default_params(res, "Dambreak"))
    orc = ob.Oracle(ob.default_params(res, "Dambreak"), np.zeros((1, 3), np.float32), boundary_seed=0)
    assert np.array_equal(w, orc.table(0)) and np.array_equal(g, orc.table(1))
    assert np.array_equal(c, orc.kernel_consts()[:3])
    assert w[10000] == 0 and g[10000] == 0 and g[0] == 0 and w[0] == c[0]
    assert np.all(np.diff(w[:10000]) <= 0) and np.all(g[1:10000] <= 0)  # W decreasing, spiky gradient attractive sign
    orc.close()


@pytest.mark.parametrize("seed", [0, 1, 12345])
def test_boundary_particles_match_oracle(sf, ob, seed):
    from simplefluid_b200 import binding
    p, po = sf.default_params(24, "Dambreak"), ob.default_params(24, "Dambreak")
    orc = ob.Oracle(po, np.zeros((1, 3), np.float32), boundary_seed=seed)
    for wall in range(6):
        b = binding.boundary_generate(p, seed, wall)
        assert b.shape == (243, 3)  # 9 x 9 x 3 for every resolution (A.4)
        assert np.array_equal(b, orc.boundary(wall))
        axis, upper = wall // 2, wall & 1
        assert np.all(b[:, axis] > 1.0) if upper else np.all(b[:, axis] < -1.0)  # patches hang outside the box
    orc.close()


def test_no_cpu_fallback(sf):
    """Without a usable B200 the product must refuse loudly, never compute on the host."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(sf.SFError) as e:
        sf.SPHSolver(sf.default_params(24, "Dambreak"))
    assert "no CPU fallback" in str(e.value)


def test_pinned_host_alloc_without_gpu(sf):
    """sf_host_alloc (page-locked buffers for sf_step_host) reports an error instead of handing out pageable memory."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(sf.SFError):
        sf.PinnedArray((16, 3))
    assert sf.library().sf_host_free(None) == 0  # freeing nothing is fine


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under simplefluid_b200/ or include/ may mention it."""
    bad = []
    for base in ("simplefluid_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    text = open(os.path.join(dirpath, fn), errors="ignore").read()
                    if re.search(r"sf_oracle|oracle_binding|libsf_oracle|sfo_", text):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_no_exception_crosses_the_c_boundary():
    """include/sf_b200.h promises that nothing throws across the boundary: every status-returning entry point with a body
    of its own is a function-try-block that maps bad_alloc / length_error to SF_ERR_OOM (csrc/sf_api.cu, SF_NOTHROW)."""
    text = open(os.path.join(ROOT, "simplefluid_b200", "csrc", "sf_api.cu")).read()
    text = text[text.index('extern "C" {'):]
    lines = text.split("\n")
    entries, guarded = [], []
    for i, line in enumerate(lines):
        m = re.match(r"^int (sf_\w+)\(.*\)$", line)
        if m:
            entries.append(m.group(1))
            if lines[i + 1] == "try {":
                guarded.append(m.group(1))
    assert len(entries) > 40
    assert entries == guarded, sorted(set(entries) - set(guarded))
    assert text.count("SF_NOTHROW(") == len(entries) + 1  # + the definition


def test_every_entry_point_rejects_a_null_solver(sf):
    """The reference Q_ASSERTs on a null solver (Source/Simulator.cpp:35-36); the C-ABI returns SF_ERR_INVALID from every
    entry point that takes one -- before touching CUDA, so this runs without a GPU."""
    hdr = open(os.path.join(ROOT, "include", "sf_b200.h")).read()
    entries = re.findall(r"^int\s+(sf_\w+)\(sf_solver\* s([^)]*)\)", hdr, re.M)
    assert len(entries) >= 40
    L = C.CDLL(sf.library_path())
    for name, rest in entries:
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = None
        args = []
        for a in [a.strip() for a in rest.split(",") if a.strip()]:
            args.append(C.c_double(0.0) if a.startswith("double") else (C.c_float(0.0) if a.startswith("float ") else C.c_void_p(0)))
        assert fn(None, *args) == -1, name


def test_public_header_is_plain_c():
    """include/sf_b200.h is the C-ABI: it must compile as C99 with no C++ or CUDA in sight."""
    import subprocess
    import tempfile
    src = '#include "sf_b200.h"\nint main(void) { sf_params p; return sf_params_default(&p) == SF_OK ? 0 : (int)sizeof(sf_solver*); }\n'
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.c")
        open(path, "w").write(src)
        r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), path],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_cpp_facade_headers_compile():
    """simplefluid_b200/host/ (SPHSolver, QtSPHSolver, SceneManager, Simulator) is header-only C++17 over the C-ABI."""
    import subprocess
    import tempfile
    src = ('#include "simplefluid_b200/host/Simulator.h"\n'
           'int main() { auto p = std::make_shared<SPHParameters<float>>(); p->kernelRadius = 2.0f / 24.0f; p->updateParams();\n'
           '  SceneManager sm(p); Vec_Vec3<float> x, v; p->scene = SimulationScenes::CubeDrop; sm.setupScene(x, v);\n'
           '  return x.size() == 13824 && v.size() == x.size() ? 0 : 1; }\n')
    with tempfile.TemporaryDirectory() as d:
        path, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(path, "w").write(src)
        lib = os.path.join(ROOT, "simplefluid_b200", "lib")
        r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", ROOT, path, "-o", exe, "-L", lib, "-lsf_b200",
                            f"-Wl,-rpath,{lib}", "-lpthread"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        # scene generation is host-side: runs without a GPU and reproduces the reference's CubeDrop count (Captured/2.png)
        assert subprocess.run([exe]).returncode == 0


# ---- checkpoint files: every count is validated against the file size before anything is allocated ------------
def _checkpoint_bytes(sf, n=5, walls=(2, 0, 0, 0, 0, 1), part=0, parts=1, n_global=None, ids=None, magic=b"SFCKPT2\0"):
    import struct
    p = sf.default_params(24, "Dambreak")
    has_ids = ids is not None
    n_global = n if n_global is None else n_global
    hdr = struct.pack("<8sII6IfIIII4xQ", magic, C.sizeof(p), n, *walls, 0.25, 1, 1 if has_ids else 0, part, parts, n_global)
    assert len(hdr) == 72
    body = bytes(p)
    for w in walls:
        body += np.zeros((w, 3), np.float32).tobytes()
    x = np.linspace(-0.5, 0.5, 3 * n, dtype=np.float32).reshape(n, 3)
    body += x.tobytes() + np.zeros((n, 3), np.float32).tobytes()
    if has_ids:
        body += np.asarray(ids, np.uint32).tobytes()
    return hdr + body


def _read_checkpoint(sf, path):
    L = sf.library()
    h, t = C.c_void_p(), C.c_float(0)
    rc = L.sf_checkpoint_read(str(path).encode(), 0, C.byref(h), C.byref(t))
    msg = (L.sf_last_error(None) or b"").decode()
    if rc == 0:
        L.sf_destroy(h)
    return rc, msg


def test_checkpoint_files_are_validated_before_use(sf, tmp_path):
    import struct
    good = _checkpoint_bytes(sf)
    f = tmp_path / "good.ckpt"
    f.write_bytes(good)
    rc, msg = _read_checkpoint(sf, f)
    # without a GPU a VALID file gets as far as sf_create (no CPU fallback); with one it restores
    assert rc in (0, -2), (rc, msg)
    cases = {
        "truncated": good[:-7],
        "trailing": good + b"\0" * 4,
        "bad magic": _checkpoint_bytes(sf, magic=b"SFCKPT1\0"),
        "huge n": good[:12] + struct.pack("<I", 0xFFFFFFF0) + good[16:],
        "huge wall count": good[:16] + struct.pack("<I", 0x7FFFFFFF) + good[20:],
        "part of a slab checkpoint as a whole file": _checkpoint_bytes(sf, n=2, parts=2, n_global=4, ids=[0, 1]),
        "header only": good[:72],
        "empty": b"",
    }
    for name, data in cases.items():
        g = tmp_path / "bad.ckpt"
        g.write_bytes(data)
        rc, msg = _read_checkpoint(sf, g)
        assert rc == -1, (name, rc, msg)
    assert _read_checkpoint(sf, tmp_path / "missing.ckpt")[0] == -1


def test_slab_checkpoint_parts_must_partition_the_ids(sf, tmp_path):
    base = tmp_path / "run.ckpt"
    (tmp_path / "run.ckpt.0").write_bytes(_checkpoint_bytes(sf, n=3, part=0, parts=2, n_global=5, ids=[0, 2, 4]))
    (tmp_path / "run.ckpt.1").write_bytes(_checkpoint_bytes(sf, n=2, part=1, parts=2, n_global=5, ids=[1, 3]))
    rc, msg = _read_checkpoint(sf, base)
    assert rc in (0, -2), (rc, msg)  # consistent parts: parsed, then sf_create (GPU) or SF_ERR_CUDA (none)
    (tmp_path / "run.ckpt.1").write_bytes(_checkpoint_bytes(sf, n=2, part=1, parts=2, n_global=5, ids=[1, 2]))  # id 2 twice, 3 missing
    assert _read_checkpoint(sf, base)[0] == -1
    (tmp_path / "run.ckpt.1").write_bytes(_checkpoint_bytes(sf, n=2, part=1, parts=2, n_global=5, ids=[1, 9]))  # id out of range
    assert _read_checkpoint(sf, base)[0] == -1
    os.remove(tmp_path / "run.ckpt.1")
    assert _read_checkpoint(sf, base)[0] == -1  # a part is missing
