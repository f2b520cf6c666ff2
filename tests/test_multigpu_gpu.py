"""Multi-GPU parity (needs >= 2 B200s; skipped on a single-GPU box): the z-slab run, gathered by global particle id,
must be bit-identical to a single-GPU run of the same scene (tools/mgpu_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,scene,res,steps,env,mode", [
    (2, "Dambreak", 32, 300, {}, "resident"),                       # y-slabs (auto), migration, moving cut planes
    (4, "DoubleDambreak", 48, 120, {}, "resident"),                 # z-slabs (auto)
    (2, "Dambreak", 40, 800, {"SF_SLAB_TIGHT": "1"}, "resident"),   # no capacity headroom: on-demand growth paths
    (2, "SphereDrop", 32, 200, {"SF_SLAB_AXIS": "y"}, "resident"),  # forced axis
    (2, "Dambreak", 32, 200, {}, "host_owned"),                     # sf_step_host_owned: owned particles live on the host
    (2, "Dambreak", 32, 200, {"SF_SLAB_TIGHT": "1"}, "host_owned"),
    (2, "DoubleDambreak", 40, 160, {}, "checkpoint"),               # per-rank checkpoint parts, restored on 2 ranks and on 1
    (2, "Dambreak", 32, 250, {"SF_TEST_PARAMS": "bCorrectDensity=1"}, "resident"),  # Shepard pass: four ghost layers per side
    (2, "CubeDrop", 24, 400, {"SF_TEST_PARAMS": "bCorrectDensity=1,bUseAttractivePressure=1", "SF_SLAB_AXIS": "y"}, "resident"),
])
def test_slabs_bit_identical_to_single_gpu(sf, world, scene, res, steps, env, mode):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "mgpu_check.py"), scene, str(res), str(steps), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, **env})
    assert out.returncode == 0 and "MGPU OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
