"""Multi-GPU slab check (run under torchrun; optional 4th argument: resident | host_owned | checkpoint): every rank advances its slab; rank 0 gathers the owned particles
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
    mode = sys.argv[4] if len(sys.argv) > 4 else "resident"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [binding.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    over = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("SF_TEST_PARAMS", "").split(",") if kv)}  # e.g. bCorrectDensity=1
    p = sf.default_params(res, scene, **over)
    pos = sf.scene_generate(p)
    g = sf.SPHSolver(p, device=local)
    g.commInit(rank, world, uid[0])
    g.setParticlesGlobal(pos)
    g.generateBoundaryParticles(0)
    g.makeReady()
    dts = []
    infos = [g.slabInfo()]
    t0 = time.time()
    if mode == "host_owned":
        # the host holds the owned particles between substeps (sf_step_host_owned): what bench.py's e2e leg does
        cap = int(g.localSlots() * 2) + 4096
        hi, hx, hv = np.empty(cap, np.uint32), np.empty((cap, 3), np.float32), np.empty((cap, 3), np.float32)
        m = g.downloadOwnedInto(hi, hx, hv)
        for k in range(steps):
            if k % 7 == 3:  # the host may hand the particles back in any order
                perm = np.random.default_rng(k).permutation(m)
                hi[:m], hx[:m], hv[:m] = hi[:m][perm], hx[:m][perm], hv[:m][perm]
            m = g.stepHostOwned(hi, hx, hv, m)
            dts.append(g.last_dt)
            if k % max(1, steps // 6) == 0:
                infos.append(g.slabInfo())
    elif mode == "checkpoint":
        # half of the substeps, checkpoint (one part per rank), restore on the same ranks, the other half
        import tempfile
        d = [tempfile.mkdtemp(prefix="sfckpt") if rank == 0 else None]
        dist.broadcast_object_list(d, src=0)
        path = os.path.join(d[0], "run.ckpt")
        for k in range(steps // 2):
            dts.append(g.advanceFrame())
        g.checkpointWrite(path, sim_time=float(np.sum(dts)))
        g.close()
        dist.barrier()
        uid2 = [binding.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid2, src=0)
        g, t_restored = sf.SPHSolver.fromCheckpointSlab(path, local, rank, world, uid2[0])
        assert abs(t_restored - float(np.sum(dts))) < 1e-6
        infos.append(g.slabInfo())
        for k in range(steps - steps // 2):
            dts.append(g.advanceFrame())
        if rank == 0:  # the same parts restored on ONE GPU continue identically, too
            one, _ = sf.SPHSolver.fromCheckpoint(path, device=local)
            for k in range(steps - steps // 2):
                one.advanceFrame()
            single_from_parts = (one.getParticles(), one.getVelocity())
            one.close()
    else:
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
            if mode == "checkpoint":
                e1 = np.array_equal(single_from_parts[0], rx) and np.array_equal(single_from_parts[1], rv)
                print(f"slab checkpoint restored on one GPU continues bit-identically: {e1}")
                ok = ok and e1
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
