#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -12 gpurun_out/r2v_pytest.log
for w in cube_1m sphere_16m; do
  timeout 600 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_bench_$w.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2v_bench_$w.log").read().strip().splitlines()[-1])
    print("$w", d["config"]["particles_total"], "dev %.3e (%.3f ms) rest %.3e (%.3f ms)"%(d["value"],d["ms_per_step"],d["value_at_rest"],d["ms_per_step_at_rest"]), d["flow"]["developed"], "parity", d.get("parity_vs_reference_binary"))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/r2v_bench_$w.log").read()[-1500:])
PY
done
timeout 300 python tools/exp_bench.py 203 1500 40
