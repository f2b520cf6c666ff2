"""GPU test of the C++ host facade (simplefluid_b200/host/: Simulator / QtSPHSolver / SceneManager over the
C-ABI) through the headless driver sf_sim, against the same run made through the ctypes binding."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sf_sim_matches_binding(sf, tmp_path):
    exe = os.path.join(os.path.dirname(sf.library_path()), "sf_sim")
    if not os.path.exists(exe):
        sf.build_library()
    prefix = str(tmp_path / "frame")
    out = subprocess.run([exe, "--scene", "DoubleDambreak", "--resolution", "16", "--stop-time", "0.1", "--dump-prefix", prefix,
                          "--seed", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    info = json.loads(out.stdout.strip().splitlines()[-1])
    # the same three frames through the Python mirror of the C-ABI
    p = sf.default_params(16, "DoubleDambreak")
    pos = sf.scene_generate(p)
    assert info["particles"] == len(pos) and info["frames"] == 3  # 0.1 s at 1/30 s per frame
    gpu = sf.SPHSolver(p)
    gpu.setParticles(pos)
    gpu.generateBoundaryParticles(0)
    gpu.makeReady()
    sim_time = np.float32(0)
    for frame in range(1, 4):
        t, _ = gpu.advanceFrameTime(0.0333333333)
        sim_time = np.float32(sim_time + np.float32(t))
        with open(f"{prefix}.{frame:04d}.bin", "rb") as f:
            magic, n, tfile = f.read(4), *struct.unpack("<If", f.read(8))
            x = np.frombuffer(f.read(), np.float32).reshape(-1, 3)
        assert magic == b"SFF1" and n == len(pos) and np.float32(tfile) == sim_time
        assert np.array_equal(x, gpu.getParticles())
    assert abs(info["sim_time"] - float(sim_time)) < 1e-6
    gpu.close()


def _frames(prefix, count):
    out = []
    for frame in range(1, count + 1):
        with open(f"{prefix}.{frame:04d}.bin", "rb") as f:
            f.read(12)
            out.append(np.frombuffer(f.read(), np.float32).reshape(-1, 3))
    return out


def test_stop_and_start_again_continues_the_flow(sf, tmp_path):
    """Simulator::stop() then startSimulation() (Source/Simulator.cpp:22-30,66-69): makeReady() runs again on the live
    particle set, so the frames after the pause equal those of an uninterrupted run (round-1 advisor finding: the
    facade used to re-upload the last host-synced state)."""
    exe = os.path.join(os.path.dirname(sf.library_path()), "sf_sim")
    base = [exe, "--scene", "Dambreak", "--resolution", "16", "--stop-time", "0.16", "--seed", "0"]
    a = subprocess.run(base + ["--dump-prefix", str(tmp_path / "a")], capture_output=True, text=True, timeout=300)
    b = subprocess.run(base + ["--dump-prefix", str(tmp_path / "b"), "--pause-frame", "2"], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    ia, ib = json.loads(a.stdout.strip().splitlines()[-1]), json.loads(b.stdout.strip().splitlines()[-1])
    assert ia["frames"] == ib["frames"] == 5 and abs(ia["sim_time"] - ib["sim_time"]) < 1e-7
    fa, fb = _frames(str(tmp_path / "a"), 5), _frames(str(tmp_path / "b"), 5)
    for k in range(5):
        assert np.array_equal(fa[k], fb[k]), f"frame {k + 1} differs after the pause"
    assert not np.array_equal(fa[4], fa[1])


@pytest.mark.parametrize("pause", [0, 1])
def test_reference_simulator_unmodified_on_the_facade(sf, tmp_path, pause):
    """/root/reference/Source/Simulator.cpp + SceneManager.cpp, compiled unmodified against the host facade
    (oracle/_ref/ref_simulator, built by oracle/Makefile `callers`), driven like MainWindow drives it: the "Position"
    array the renderer would upload equals the binding's positions after the same three frames -- also when the
    run is stopped after the first frame and started again."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_simulator")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_simulator not built (needs /root/reference at build time)")
    out_bin = str(tmp_path / "ref.bin")
    out = subprocess.run([exe, "2", "16", "0.1", out_bin] + (["1"] if pause else []), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    with open(out_bin, "rb") as f:
        n, frames, changed = struct.unpack("<III", f.read(12))
        x = np.frombuffer(f.read(), np.float32).reshape(-1, 3)
    p = sf.default_params(16, "Dambreak")
    pos = sf.scene_generate(p)
    gpu = sf.SPHSolver(p)
    gpu.setParticles(pos)
    gpu.makeReady()  # default walls: seed 0, what the facade's makeReady does too
    substeps = 0
    for _ in range(3):
        _, k = gpu.advanceFrameTime(0.0333333333)
        substeps += k
    assert n == len(pos) and frames == 3 and changed == substeps + 1  # one particleChanged per substep (Simulator.cpp:50) + setupScene's
    assert np.array_equal(x, gpu.getParticles())
    gpu.close()
