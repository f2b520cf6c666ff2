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
