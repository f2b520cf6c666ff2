"""Slab throughput probe (torchrun): python -m torch.distributed.run --nproc-per-node N tools/mgpu_bench.py Dambreak <res> <steps>"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplefluid_b200 as sf  # noqa: E402
from simplefluid_b200 import binding  # noqa: E402

scene, res, steps = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = [binding.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
p = sf.default_params(res, scene)
pos = sf.scene_generate(p)
g = sf.SPHSolver(p, device=local)
g.commInit(rank, world, uid[0])
g.setParticlesGlobal(pos)
g.makeReady()
def _die(exc):
    # never hang in NCCL / torch teardown while the other ranks wait in a collective
    import traceback
    traceback.print_exception(exc)
    sys.stderr.flush()
    os._exit(1)


sys.excepthook = lambda t, e, tb: _die(e)
g.advanceSteps(5)
g.synchronize()
dist.barrier()
g.profileEnable(True)
g.profileReset()
t0 = time.perf_counter()
g.timerStart()
g.advanceSteps(steps)
ms = g.timerStop()
wall = time.perf_counter() - t0
prof = g.profile()
info = g.slabInfo()
out = [None] * world if rank == 0 else None
dist.gather_object((ms, wall, info, {k: v[0] / steps for k, v in prof.items() if v[1]}), out, dst=0)
if rank == 0:
    worst = max(o[0] for o in out)
    print(f"{scene} res {res} N={len(pos)} world={world}: {worst / steps:.3f} ms/step -> {len(pos) * steps / worst * 1e3:.3e} particle-steps/s")
    for r, o in enumerate(out):
        print(f"  rank {r}: {o[0] / steps:.3f} ms/step (wall {o[1] / steps * 1e3:.3f}) slab {o[2]}  " +
              " ".join(f"{k[2:]}={v:.3f}" for k, v in o[3].items() if v > 0.02))
g.close()
dist.barrier()
dist.destroy_process_group()
