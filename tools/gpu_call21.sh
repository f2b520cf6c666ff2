#!/bin/bash
mkdir -p gpurun_out
for w in cube_1m sphere_16m dambreak_default doubledambreak_8m; do
  timeout 600 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_$w.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2u_bench_$w.log").read().strip().splitlines()[-1])
    print("$w", d["config"]["particles_total"], "dev %.3e (%.3f ms) rest %.3e (%.3f ms)"%(d["value"],d["ms_per_step"],d["value_at_rest"],d["ms_per_step_at_rest"]), d["flow"]["developed"], "parity", d.get("parity_vs_reference_binary"), "e2e %.3e"%d["e2e"]["value"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/r2u_bench_$w.log").read()[-1500:])
PY
done
