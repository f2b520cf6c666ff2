#!/usr/bin/env python
"""bench.py -- particle-steps/s of the SPH substep on N B200s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--settle S] [--workload NAME] [--impl b200|reference]
                    [--no-verify] [--no-cpu-baseline]

One "step" = one reference substep (advanceFrame, Source/Simulator.cpp:49) over the whole particle
set.  Workloads (BASELINE.json configs; default = the weak-scaling unit of configs[4]):

    dambreak_8m_per_gpu  Dambreak, ~8.09M particles per GPU (res 203 at N=1 ... res 404 = 64.2M at N=8); weak scaling
    doubledambreak_8m    configs[2]  DoubleDambreak res 161, 8,028,160 particles on 1/2/4/8 GPUs; STRONG scaling
    cube_1m              configs[1]  CubeDrop res 100, 1,000,000 particles
    sphere_16m           configs[3]  SphereDrop res 313, 16,054,752 particles
    dambreak_default     configs[0]  Dambreak res 24, 11,979 particles

Two timed regions of K substeps each, both device-resident and timed with CUDA events on the solver's stream (max over
ranks): `value_at_rest` right after the W warm-up substeps (the initial lattice: every particle has <= 32 neighbours),
then S settle substeps (untimed; the column collapses, the flow develops to rest density with ragged cells), then
`value` on that DEVELOPED state -- the headline.  The per-kernel numbers (`roofline`, `kernel_share`) belong to the
developed region.  `e2e` = the same metric through the host-buffer C-ABI calls (pinned host buffers, H2D + D2H inside
the timed region); `cpu_baseline` = the CPU oracle ("port": line-faithful transcription, the reference cannot be built
-- BASELINE.md section 2) on this box's host cores on a bounded sample of the SAME scene and resolution; at N > 1
`parity_vs_single_gpu` = the slab run gathered by particle id and compared bit for bit with a single-GPU run of the
same substeps on rank 0; at N = 1 `parity_vs_reference_binary` = the first substeps of the run against the checksums of
the reference binary's own outputs for the same scene (tests/golden/exe_fullsize_checksums.json).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic HBM bytes per particle-step (DESIGN.md "Kernels")
PROFILE_EVERY = 8
ALGO_BYTES = {"step": 156.0, "k_density": 16.0, "k_force": 40.0, "k_visc_integrate": 52.0, "sort_reorder": 48.0}

WORKLOADS = {
    # name: (scene, {n_gpus: resolution}, scaling)
    "dambreak_8m_per_gpu": ("Dambreak", {1: 203, 2: 255, 4: 321, 8: 404}, "weak"),
    "doubledambreak_8m": ("DoubleDambreak", {1: 161, 2: 161, 4: 161, 8: 161}, "strong"),
    "cube_1m": ("CubeDrop", {1: 100}, "weak"),
    "sphere_16m": ("SphereDrop", {1: 313}, "weak"),
    "dambreak_default": ("Dambreak", {1: 24}, "weak"),
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md


def workload_config(workload, n_gpus):
    """The `config` object both arms print (identical for --impl b200 and --impl reference)."""
    scene, res_by_n, scaling = WORKLOADS[workload]
    if n_gpus not in res_by_n:
        raise SystemExit(f"workload {workload} is not defined for {n_gpus} GPUs")
    return scene, res_by_n[n_gpus], scaling


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = sorted(x for x in sm if x > 0.5 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return rank, world, local, None


def barrier(dist):
    if dist is not None:
        dist.barrier()


def reduce_ranks(dist, x, local, op="max"):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def bind_to_gpu_numa_node(local):
    """Pinned host buffers should live on the NUMA node the GPU hangs off: bind this rank's CPU affinity to the GPU's
    local CPUs (nvidia-smi topology) before allocating them.  Best effort; a no-op when the topology is flat."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(local)], capture_output=True, text=True, timeout=10).stdout
        # e.g. "NUMA IDs of closest CPU: 0" ; fall back to the affinity mask of `nvidia-smi topo -m` when absent
        node = None
        for tok in out.replace(":", " ").split():
            if tok.isdigit():
                node = int(tok)
        if node is None:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return node
    except Exception:
        return None


def cpu_oracle_run(scene, res, steps, warmup, threads=0, budget_s=None):
    """Times the CPU oracle (the reference's algorithm restated, oracle/) on this box's host cores.  With a time
    budget the run stops early once it is spent (at least 2 timed substeps); returns the substeps actually timed."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    p = ob.default_params(res, scene)
    pos = ob.scene(p)
    orc = ob.Oracle(p, pos, boundary_seed=0, threads=threads)
    t_all = time.perf_counter()
    for _ in range(warmup):
        orc.advance()
        if budget_s is not None and time.perf_counter() - t_all > 0.25 * budget_s:
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        orc.advance()
        done += 1
        if budget_s is not None and done >= 2 and time.perf_counter() - t_all > budget_s:
            break
    dt = time.perf_counter() - t0
    n = len(pos)
    orc.close()
    return n * done / dt, n, dt, done


def base_config(workload, scene, res, n_total, n_gpus):
    """Identical in both arms (--impl b200 / reference): what the workload IS.  How the GPU arm ran it (slab axis, ghost
    fraction, settle time ...) goes into the top-level `run` object."""
    if n_total is None:  # the reference arm at N > 1 did not generate the large scene: count it through the C-ABI (host side only)
        import simplefluid_b200 as sf
        import ctypes as C
        p = sf.default_params(res, scene)
        cnt = C.c_uint64(0)
        sf.library().sf_scene_generate(C.byref(p), sf.SCENES[scene], None, 0, C.byref(cnt))
        n_total = cnt.value
    return {"workload": workload, "scene": scene, "resolution": res, "particles_total": int(n_total),
            "particles_per_gpu": int(n_total // n_gpus), "grid_cells": int(res) ** 3, "boundary_seed": 0,
            "parallelism": f"slab x{n_gpus} (one process per GPU)" if n_gpus > 1 else "single",
            "l2": "inputs_exceed_l2" if (n_total // n_gpus) * 100 > 126e6 else "state_fits_l2_not_flushed"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself cannot be
    built (its solver lives in the un-vendored Banana library; no Qt/TBB here), so this is the oracle
    port, with all host threads, on the SAME scene and resolution as the GPU arm at this N, from the initial
    lattice (its cheapest state: the developed flow the GPU arm's `value` is timed on has more neighbours per
    particle)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, res, scaling = workload_config(args.workload, args.gpus)
    # The CPU arm runs on ONE host whatever N is.  At N = 1 it runs the GPU arm's very configuration; at N > 1 a weak-scaling
    # workload would be N times larger (64 M particles at N = 8: 8 s per substep), so it runs the N = 1 size of the same
    # workload and says so -- the metric is per particle-step and the CPU cost per particle does not fall with size.
    cpu_res = res if scaling == "strong" else workload_config(args.workload, 1)[1]
    cores = len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    value, n, secs, done = cpu_oracle_run(scene, cpu_res, args.steps, args.warmup, budget_s=150.0)
    sample = (f"{scene} res {cpu_res} ({n} particles" + (", the GPU arm's configuration" if cpu_res == res else
              f", the N = 1 size of this workload; the GPU arm at {args.gpus} GPUs runs res {res}") +
              f") x {done} substeps from the initial lattice, OpenMP {cores} threads, serial cell insertion as in the reference")
    config = base_config(args.workload, scene, res, n if cpu_res == res else None, args.gpus)
    if cpu_res != res:
        config["reference_arm_resolution"] = cpu_res
        config["reference_arm_particles"] = int(n)
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": secs / done * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def timed_region(gpu, dist, local, steps, profile):
    """K substeps between two in-stream CUDA events, a barrier and a synchronize on both sides; ms = max over ranks."""
    if profile:
        gpu.profileEnable(True, every=PROFILE_EVERY)  # per-kernel events on every 8th substep, graph replay otherwise
        gpu.profileReset()
    launches0 = gpu.launchCount()
    gpu.synchronize()
    barrier(dist)
    gpu.timerStart()
    gpu.advanceSteps(steps)
    ms = gpu.timerStop()
    gpu.synchronize()
    barrier(dist)
    launches = gpu.launchCount() - launches0
    prof = gpu.profile() if profile else None
    if profile:
        gpu.profileEnable(False)
    return reduce_ranks(dist, ms, local), launches, prof


def gather_diag(gpu, dist, local):
    d = gpu.diagnostics()
    out = {
        "nbr_mean": reduce_ranks(dist, float(d["nbr_sum"]), local, "sum") / max(reduce_ranks(dist, float(d["particles"]), local, "sum"), 1.0),
        "nbr_max": int(reduce_ranks(dist, float(d["nbr_max"]), local)),
        "fallback_bricks": int(reduce_ranks(dist, float(d["fallback_bricks"]), local, "sum")),
        "fallback_particles": int(reduce_ranks(dist, float(d["fallback_particles"]), local, "sum")),
        "particles_without_list": int(reduce_ranks(dist, float(d["particles_without_list"]), local, "sum")),
        "bricks": int(reduce_ranks(dist, float(d["bricks"]), local, "sum")),
    }
    out["nbr_mean"] = round(out["nbr_mean"], 2)
    return out


def verify_vs_single_gpu(sf, gpu, dist, rank, local, p, pos, total_steps):
    """Owned particles of every rank, gathered by global id on rank 0, against a single-GPU run of the same number
    of substeps from the same initial set on rank 0's device.  Bit-identical or False."""
    ids, x, v = gpu.downloadOwned()
    gathered = [None] * dist.get_world_size() if rank == 0 else None
    dist.gather_object((ids, x, v), gathered, dst=0)
    ok = None
    if rank == 0:
        allids = np.concatenate([a[0] for a in gathered])
        allx = np.concatenate([a[1] for a in gathered])
        allv = np.concatenate([a[2] for a in gathered])
        del gathered
        order = np.argsort(allids, kind="stable")
        ok = len(allids) == len(pos) and bool(np.array_equal(allids[order], np.arange(len(pos), dtype=np.uint32)))
        ref = sf.SPHSolver(p, device=local)
        ref.setParticles(pos)
        ref.generateBoundaryParticles(0)
        ref.makeReady()
        ref.advanceSteps(total_steps)
        rx, rv = ref.getParticles(), ref.getVelocity()
        ref.close()
        ok = bool(ok and np.array_equal(allx[order], rx) and np.array_equal(allv[order], rv))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    return flag[0]


def verify_vs_reference_binary(gpu, scene, res):
    """N = 1: the first substeps of this very run against the checksums of the REFERENCE BINARY's own outputs for the same
    scene and resolution (tests/golden/exe_fullsize_checksums.json; tests/golden/make_exe_fullsize.py produced them by
    executing Prebuild/SimpleFluid.exe's compiled step).  Returns (flag or None when no fixture covers the scene, substeps used)."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "tests", "golden", "exe_fullsize_checksums.json")) as f:
            cases = json.load(f)["cases"]
    except Exception:
        return None, 0
    rec = next((c for c in cases.values() if c["scene"] == scene and int(c["resolution"]) == int(res)), None)
    if rec is None:
        return None, 0
    digest = lambda a: hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()
    ok = True
    for k, want in enumerate(rec["steps"]):
        ok = ok and float(np.float32(gpu.advanceFrame())) == float(np.float32(rec["dts"][k]))
        got = {"cell": gpu.cellIndex(), "rho": gpu.density(), "x": gpu.getParticles(), "v": gpu.getVelocity()}
        ok = ok and all(digest(got[f]) == want[f] for f in got)
    return bool(ok), len(rec["steps"])


def run_b200(args):
    import simplefluid_b200 as sf

    rank, world, local, dist = dist_setup(args.gpus)
    n_gpus = max(world, 1)
    scene, res, scaling = workload_config(args.workload, n_gpus)
    multi = n_gpus > 1
    numa = bind_to_gpu_numa_node(local) if multi else None
    # weak scaling: ~8.09M particles per GPU; at N > 1 the ONE scene of N x 8M particles is cut into slabs
    p = sf.default_params(res, scene)
    pos = sf.scene_generate(p)
    n_total = len(pos)

    gpu = sf.SPHSolver(p, device=local)
    if multi:
        from simplefluid_b200 import binding
        uid = [binding.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gpu.commInit(rank, world, uid[0])
        gpu.setParticlesGlobal(pos)
    else:
        gpu.setParticles(pos)
    gpu.generateBoundaryParticles(0)
    gpu.makeReady()
    total_particles = float(n_total)

    # ---- N = 1: the first substeps against the reference binary's own outputs (part of the warm-up) -------------
    ref_parity, used = (None, 0)
    if not multi and not args.no_verify:
        ref_parity, used = verify_vs_reference_binary(gpu, scene, res)

    # ---- device-resident throughput: at rest, then on the developed flow ------------------------
    gpu.advanceSteps(max(args.warmup - used, 0))
    gpu.synchronize()
    barrier(dist)
    ms_rest, _, prof_rest = timed_region(gpu, dist, local, args.steps, profile=True)
    diag_rest = gather_diag(gpu, dist, local)
    value_rest = total_particles * args.steps / (ms_rest * 1e-3)
    t_settle = time.perf_counter()
    done = 0
    while done < args.settle:  # in batches: the launch queue of a slab run is driven by the host
        k = min(250, args.settle - done)
        gpu.advanceSteps(k)
        gpu.synchronize()
        done += k
    t_settle = time.perf_counter() - t_settle
    barrier(dist)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ms, launches, prof = timed_region(gpu, dist, local, args.steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    diag = gather_diag(gpu, dist, local)
    value = total_particles * args.steps / (ms * 1e-3)
    slab_axis = None
    n = n_total
    if multi:
        zb, ze, n_own, n_ghost = gpu.slabInfo()
        n = n_own
        slab_axis = gpu.slabAxis()
        ghost_frac = reduce_ranks(dist, float(n_ghost), local, "sum") / total_particles
        own_max = reduce_ranks(dist, float(n_own), local)
        thick_min = -reduce_ranks(dist, -float(ze - zb), local)

    # ---- parity of the slab run against a single-GPU run of the same substeps ---------------------
    parity = None
    if multi and not args.no_verify:
        parity = verify_vs_single_gpu(sf, gpu, dist, rank, local, p, pos, args.warmup + 2 * args.steps + args.settle)

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------
    e2e_steps = max(3, min(args.steps, 20))
    if not multi:
        hx, hv = sf.PinnedArray((n_total, 3)), sf.PinnedArray((n_total, 3))
        hx.array[:] = gpu.getParticles()
        hv.array[:] = gpu.getVelocity()
        for _ in range(2):
            gpu.stepHost(hx.array, hv.array)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            gpu.stepHost(hx.array, hv.array)
        e2e_s = time.perf_counter() - t0
        h2d = d2h = 24 * n_total
        e2e_api = "sf_step_host (pinned host buffers, upload + substep + download per step, developed state)"
    else:
        # every substep: this rank's OWNED particles host -> device (24 B each), one substep incl. the halo exchange,
        # owned particles device -> host; copies and kernels overlap as in sf_step_host
        cap = int(gpu.localSlots() * 1.5) + (1 << 18)
        hx, hv, hi = sf.PinnedArray((cap, 3)), sf.PinnedArray((cap, 3)), sf.PinnedArray((cap,), np.uint32)
        m = gpu.downloadOwnedInto(hi.array, hx.array, hv.array)
        moved_up = moved_down = 0
        for it in range(2 + e2e_steps):
            if it == 2:
                barrier(dist)
                t0 = time.perf_counter()
                moved_up = moved_down = 0
            m_up = m
            m = gpu.stepHostOwned(hi.array, hx.array, hv.array, m)
            moved_up += 28 * m_up
            moved_down += 28 * m
        gpu.synchronize()
        e2e_s = reduce_ranks(dist, time.perf_counter() - t0, local)
        h2d = int(reduce_ranks(dist, float(moved_up), local, "sum") / e2e_steps)
        d2h = int(reduce_ranks(dist, float(moved_down), local, "sum") / e2e_steps)
        e2e_api = ("sf_step_host_owned (pinned host buffers on the GPU's NUMA node; every rank's OWNED particles, 28 B each way: "
                   "id + position + velocity; upload + substep incl. halo exchange + download per step, developed state)")
    e2e_value = total_particles * e2e_steps / e2e_s

    if rank != 0:
        gpu.close()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = measured_peaks()
    step_kernels = {k: v for k, v in prof.items() if v[1] and k != "k_marshal"}
    dom = max(step_kernels, key=lambda k: step_kernels[k][0])
    dom_ms = step_kernels[dom][0] / step_kernels[dom][1]
    dom_bytes = ALGO_BYTES.get(dom, 0.0) * n  # rank 0's own particles (its launch also covers the ghost layers)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    step_achieved = ALGO_BYTES["step"] * (value / n_gpus) / 1e9
    traffic, traffic_src = None, None
    try:  # ncu --set full capture of THIS workload at N = 1 on the developed state (tools/ncu_summary.py); null otherwise
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if n_gpus == 1 and tj.get("_workload") == args.workload and tj.get("_settle") == args.settle:
            traffic, traffic_src = tj.get(dom), tj.get("_source")
    except Exception:
        pass
    # per SUBSTEP: a slab substep launches the force and integrate kernels twice (edge layers, interior layers) and
    # k_begin_step twice, so totals are divided by the number of profiled substeps (= launches of the density pass)
    nprof = max(int(step_kernels.get("k_density", (0.0, 1))[1]), 1)
    rest_dom = None
    if prof_rest and prof_rest.get(dom, (0.0, 0))[1]:  # the same kernel on the rest lattice (the state round 1 measured)
        rest_ms = prof_rest[dom][0] / prof_rest[dom][1]
        rest_dom = {"avg_launch_ms": rest_ms, "achieved": dom_bytes / (rest_ms * 1e-3) / 1e9, "frac": dom_bytes / (rest_ms * 1e-3) / 1e9 / peak}
    ksum = sum(x[0] for x in step_kernels.values()) / nprof
    kernel_share = {k: round(v[0] / nprof / ksum, 4) for k, v in step_kernels.items()}
    kernel_ms = {k: round(v[0] / nprof, 4) for k, v in step_kernels.items()}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only): same scene and resolution -----------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0)) or os.cpu_count() or 1
        csteps = max(2, min(80, int(12e6 * 12 / max(n_total, 1))))  # ~10-20 s of CPU work
        cv, cn, secs, csteps = cpu_oracle_run(scene, res, csteps, 1, budget_s=40.0)
        cpu = {"value": cv, "unit": "particle-steps/s", "cores": cores, "kind": "port",
               "sample": f"{scene} res {res} ({cn} particles, same configuration) x {csteps} substeps from the initial lattice in {secs:.1f} s, OpenMP {cores} threads"}

    config = base_config(args.workload, scene, res, n_total, n_gpus)
    assert config["grid_cells"] == int(np.prod(gpu.gridDims()))
    run = {"state": f"developed flow: {args.warmup} warm-up + {args.steps} at-rest + {args.settle} settle substeps before the timed region",
           "settle_substeps": args.settle, "settle_wall_s": round(t_settle, 2)}
    if multi:
        run.update({"slab_axis": "xyz"[slab_axis], "ghost_layers_per_side": 3, "ghost_fraction": round(ghost_frac, 4),
                    "owned_max_over_mean": round(own_max / (total_particles / n_gpus), 4),
                    "slab_thickness_min_layers": int(thick_min), "numa_node_rank0": numa})
    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "value_at_rest": value_rest, "ms_per_step_at_rest": ms_rest / args.steps,
        "run": run, "flow": {"developed": diag, "at_rest": diag_rest},
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": e2e_api},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_particle": ALGO_BYTES.get(dom),
                     "avg_launch_ms": dom_ms, "timed_launches": int(step_kernels[dom][1]),
                     "timing": f"CUDA events on the solver's stream around every kernel of every {PROFILE_EVERY}th substep of the timed region",
                     "whole_step": {"achieved": step_achieved, "frac": step_achieved / peak, "algorithmic_bytes_per_particle_step": 156},
                     "state": "developed flow", "at_rest": rest_dom},
        "kernel_share": kernel_share, "kernel_ms": kernel_ms,
        "cpu_baseline": cpu,
    }
    if multi:
        line["parity_vs_single_gpu"] = parity
    else:
        line["parity_vs_reference_binary"] = ref_parity  # null: no fixture for this scene / --no-verify
    print(json.dumps(line), flush=True)
    gpu.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--settle", type=int, default=1500, help="untimed substeps between the at-rest and the developed timed region")
    ap.add_argument("--workload", default="dambreak_8m_per_gpu", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the bit-for-bit comparison with a single-GPU run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    except BaseException:
        # a failing rank must not sit in NCCL / torch teardown while the others wait in a collective
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
