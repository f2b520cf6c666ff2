"""Multi-GPU slab check (run under torchrun): every rank advances its slab; rank 0 gathers the owned particles
and compares them, by global id, with a single-GPU run of the same scene.  Expectation: bit-identical.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py Dambreak 48 60
"""
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


def main():
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
    g.generateBoundaryParticles(0)
    g.makeReady()
    dts = []
    infos = [g.slabInfo()]
    t0 = time.time()
    for k in range(steps):
        dts.append(g.advanceFrame())
        if k % max(1, steps // 6) == 0:
            infos.append(g.slabInfo())
    wall = time.time() - t0
    ids, x, v = g.downloadOwned()
    infos.append(g.slabInfo())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((ids, x, v, dts, infos), gathered, dst=0)
    ok = True
    if rank == 0:
        allids = np.concatenate([a[0] for a in gathered])
        allx = np.concatenate([a[1] for a in gathered])
        allv = np.concatenate([a[2] for a in gathered])
        for r, a in enumerate(gathered):
            print(f"rank {r}: owned {len(a[0])}  slab history {a[4]}")
            if a[3] != gathered[0][3]:
                print("  dt sequence differs from rank 0")
                ok = False
        order = np.argsort(allids, kind="stable")
        if len(allids) != len(pos) or not np.array_equal(allids[order], np.arange(len(pos), dtype=np.uint32)):
            print(f"ownership is not a partition: {len(allids)} owned of {len(pos)}, unique {len(np.unique(allids))}")
            ok = False
        # single-GPU reference on this rank's device
        ref = sf.SPHSolver(p, device=local)
        ref.setParticles(pos)
        ref.generateBoundaryParticles(0)
        ref.makeReady()
        rdts = [ref.advanceFrame() for _ in range(steps)]
        rx, rv = ref.getParticles(), ref.getVelocity()
        if ok:
            ex, ev = np.array_equal(allx[order], rx), np.array_equal(allv[order], rv)
            print(f"N={len(pos)} steps={steps} world={world}: dt equal {rdts == gathered[0][3]}  positions bit-identical {ex}  velocities {ev}"
                  f"  max|dx| {np.abs(allx[order] - rx).max():.3e}  ({wall / steps * 1e3:.2f} ms/step incl. host sync)")
            ok = ok and ex and ev and rdts == gathered[0][3]
        print("MGPU", "OK" if ok else "FAILED")
        ref.close()
    g.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
