"""The reference's own callers of the solver, compiled UNMODIFIED against the product's host facade (row (b) of
SURVEY.md section 8: the drop-in boundary).  oracle/Makefile `callers` builds /root/reference/Source/SceneManager.cpp
and /root/reference/Source/Simulator.cpp from where they lie, against simplefluid_b200/host/compat (the Banana include
paths) with the product's QtSPHSolver.h in place of Include/QtSPHSolver.h, Qt and TBB stood in by tests/stubs/.
The binaries land in oracle/_ref/ and travel to the GPU box; the GPU half is in tests/test_host_facade_gpu.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SCENE_CHECK = os.path.join(ROOT, "oracle", "_ref", "ref_scene_check")
SIMULATOR = os.path.join(ROOT, "oracle", "_ref", "ref_simulator")


def build_callers(sf):
    if not os.path.exists(sf.library_path()):
        sf.build_library()
    if os.path.isdir(REF):
        out = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "--no-print-directory", "callers"], capture_output=True, text=True)
        assert out.returncode == 0, "the reference's callers no longer compile against the host facade:\n" + out.stdout[-3000:] + out.stderr[-3000:]
    elif not os.path.exists(SCENE_CHECK):
        pytest.skip("no /root/reference here and no prebuilt oracle/_ref callers")


def test_reference_scene_manager_compiles_and_matches(sf):
    """Source/SceneManager.cpp (Vec3 arithmetic, glm::length, Vec3<int>(float,float,float), SPHParameters fields) against
    host/SPHSolver.h: every scene it fills is byte-identical to sf_scene_generate; counts are the reference's screenshots'."""
    build_callers(sf)
    out = subprocess.run([SCENE_CHECK], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict((ln.split()[0], ln.split()[1:]) for ln in out.stdout.strip().splitlines())
    assert lines["CubeDrop@24"] == ["13824", "identical"]
    assert lines["SphereDrop@24"] == ["7145", "identical"]
    assert lines["DoubleDambreak@24"] == ["23958", "identical"]
    assert lines["Dambreak@24"] == ["11979", "identical"]
    assert all(v[1] == "identical" for v in lines.values()) and len(lines) == 8


def test_reference_simulator_compiles_and_fails_loudly_without_gpu(sf):
    """Source/Simulator.cpp (doSimulation, start/stop/reset, changeScene, setupScene) against the facade.  Without a
    B200 the binary must refuse to run (no CPU fallback): exit code 3 = SPHError from sf_create."""
    build_callers(sf)
    assert os.path.exists(SIMULATOR)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: the run itself is tested in tests/test_host_facade_gpu.py")
    out = subprocess.run([SIMULATOR, "2", "12", "0.05", "/tmp/sf_ref_sim_cpu.bin"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3 and "no CPU fallback" in out.stderr, (out.returncode, out.stderr)
