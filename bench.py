#!/usr/bin/env python
"""bench.py -- particle-steps/s of the SPH substep on N B200s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

One "step" = one reference substep (advanceFrame, Source/Simulator.cpp:49) over the whole particle
set.  Workloads (BASELINE.json configs; default = the weak-scaling unit of configs[4]):

    dambreak_8m_per_gpu  Dambreak, ~8.09M particles per GPU (res 203 at N=1 ... res 404 = 64.2M at N=8)
    cube_1m              configs[1]  CubeDrop res 100, 1,000,000 particles
    doubledambreak_8m    configs[2]  DoubleDambreak res 161, 8,028,160 particles
    sphere_16m           configs[3]  SphereDrop res 313, 16,054,752 particles
    dambreak_default     configs[0]  Dambreak res 24, 11,979 particles

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events on the solver's
stream, max over ranks); `e2e` = the same metric through the host-buffer C-ABI call sf_step_host
(pinned host buffers, H2D + D2H inside the timed region); `roofline` = dominant kernel against the
measured HBM peak; `cpu_baseline` = the CPU oracle ("port": line-faithful transcription, the
reference cannot be built -- BASELINE.md section 2) timed on this box's host cores on a bounded sample.
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
    # name: (scene, {n_gpus: resolution})
    "dambreak_8m_per_gpu": ("Dambreak", {1: 203, 2: 255, 4: 321, 8: 404}),
    "cube_1m": ("CubeDrop", {1: 100}),
    "doubledambreak_8m": ("DoubleDambreak", {1: 161}),
    "sphere_16m": ("SphereDrop", {1: 313}),
    "dambreak_default": ("Dambreak", {1: 24}),
}
REFERENCE_SAMPLE_RES = {"Dambreak": 100, "CubeDrop": 100, "DoubleDambreak": 80, "SphereDrop": 124}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md


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


def max_over_ranks(dist, x, local):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, x, local):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def cpu_oracle_run(scene, res, steps, warmup, threads=0):
    """Times the CPU oracle (the reference's algorithm restated, oracle/) on this box's host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    p = ob.default_params(res, scene)
    pos = ob.scene(p)
    orc = ob.Oracle(p, pos, boundary_seed=0, threads=threads)
    for _ in range(warmup):
        orc.advance()
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.advance()
    dt = time.perf_counter() - t0
    n = len(pos)
    orc.close()
    return n * steps / dt, n, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself cannot be
    built (its solver lives in the un-vendored Banana library; no Qt/TBB here), so this is the oracle
    port, with all host threads, on a bounded sample of the same scene."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    scene, _ = WORKLOADS[workload]
    res = REFERENCE_SAMPLE_RES[scene] if workload != "dambreak_default" else 24
    cores = os.cpu_count() or 1
    value, n, secs = cpu_oracle_run(scene, res, args.steps, args.warmup)
    sample = f"{scene} res {res} ({n} particles) x {args.steps} substeps, OpenMP {cores} threads, serial cell insertion as in the reference"
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "scene": scene, "sample_resolution": res, "particles": n},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import simplefluid_b200 as sf

    rank, world, local, dist = dist_setup(args.gpus)
    n_gpus = max(world, 1)
    scene, res_by_n = WORKLOADS[args.workload]
    if n_gpus not in res_by_n:
        raise SystemExit(f"workload {args.workload} is not defined for {n_gpus} GPUs")
    multi = n_gpus > 1
    # weak scaling: ~8.09M particles per GPU; at N > 1 the ONE scene of N x 8M particles is cut into z-slabs
    res = res_by_n[n_gpus]
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
    n = n_total if not multi else gpu.slabInfo()[2]  # particles this rank owns

    # ---- device-resident throughput ------------------------------------------------------------
    gpu.advanceSteps(args.warmup)
    gpu.synchronize()
    barrier(dist)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    gpu.profileEnable(True, every=PROFILE_EVERY)  # per-kernel events on every 8th substep of the timed region, graph replay otherwise
    gpu.profileReset()
    launches0 = gpu.launchCount()
    barrier(dist)
    gpu.synchronize()
    gpu.timerStart()
    gpu.advanceSteps(args.steps)
    ms = gpu.timerStop()
    gpu.synchronize()
    barrier(dist)
    launches = gpu.launchCount() - launches0
    prof = gpu.profile()
    gpu.profileEnable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(dist, ms, local)
    total_particles = float(n_total)
    value = total_particles * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------
    import torch
    e2e_steps = max(3, min(args.steps, 20))
    if not multi:
        hx = torch.from_numpy(gpu.getParticles()).pin_memory()
        hv = torch.from_numpy(gpu.getVelocity()).pin_memory()
        for _ in range(2):
            gpu.stepHost(hx, hv)
        barrier(dist)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            gpu.stepHost(hx, hv)
        e2e_s = max_over_ranks(dist, time.perf_counter() - t0, local)
        h2d = d2h = 24 * n
        e2e_api = "sf_step_host (pinned host buffers, upload + substep + download per step)"
    else:
        # every substep: this rank's resident slab state host -> device, one substep incl. the halo exchange, device -> host
        cap = 2 * gpu.localSlots() + (1 << 20)
        hx = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
        hv = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
        hi = torch.empty((cap,), dtype=torch.int32).pin_memory()
        m = gpu.downloadLocal(hx, hv, hi)
        moved = 0
        for it in range(2 + e2e_steps):
            if it == 2:
                barrier(dist)
                t0 = time.perf_counter()
                moved = 0
            gpu.uploadLocal(hx, hv, hi, m)
            gpu.advanceFrame()
            moved += 36 * m
            m = gpu.downloadLocal(hx, hv, hi)
            moved += 36 * m
        e2e_s = max_over_ranks(dist, time.perf_counter() - t0, local)
        h2d = d2h = int(sum_over_ranks(dist, float(moved), local) / (2 * e2e_steps))
        e2e_api = "sf_upload_local + sf_advance_frame + sf_download_local (pinned host buffers, slab state of every rank, per step)"
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
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    kernel_share = {k: round(v[0] / sum(x[0] for x in step_kernels.values()), 4) for k, v in step_kernels.items()}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----------------------------------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        sres = REFERENCE_SAMPLE_RES[scene] if args.workload != "dambreak_default" else 24
        cores = os.cpu_count() or 1
        cv, cn, secs = cpu_oracle_run(scene, sres, 80, 2)
        cpu = {"value": cv, "unit": "particle-steps/s", "cores": cores, "kind": "port",
               "sample": f"{scene} res {sres} ({cn} particles) x 80 substeps in {secs:.1f} s, OpenMP {cores} threads"}

    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": scene, "resolution": res, "particles_per_gpu": int(total_particles // n_gpus),
                   "particles_total": int(total_particles), "grid_cells": int(np.prod(gpu.gridDims())),
                   "parallelism": f"zslab{n_gpus} (3-layer ghost halo, 1 NCCL exchange + 1 allreduce per substep)" if multi else "single",
                   "l2": "inputs_exceed_l2" if n * 100 > 126e6 else "state_fits_l2_not_flushed", "boundary_seed": 0},
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": e2e_api},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_particle": ALGO_BYTES.get(dom),
                     "avg_launch_ms": dom_ms, "timed_launches": int(step_kernels[dom][1]),
                     "timing": f"CUDA events on the solver's stream around every kernel of every {PROFILE_EVERY}th substep of the timed region",
                     "whole_step": {"achieved": step_achieved, "frac": step_achieved / peak, "algorithmic_bytes_per_particle_step": 156}},
        "kernel_share": kernel_share,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    gpu.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="dambreak_8m_per_gpu", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
