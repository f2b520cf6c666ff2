#!/bin/bash
# Single-GPU evidence run (round 2, final kernels): GPU suite, bench, ncu full capture (developed state), launch list,
# memcheck.  Usage on the GPU box: tools/gpu_evidence_1gpu.sh [tag]   -> gpurun_out/<tag>_*
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_bench_n1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(density|force|visc)_brick' --launch-skip 4530 --launch-count 3 -f -o gpurun_out/${tag}_dev_dam8m python tools/prof_run.py Dambreak 203 4 1510 > gpurun_out/${tag}_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 15200 -c 500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "crowded or compressed or production_list or one_substep" > gpurun_out/${tag}_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_memcheck.log
tail -4 gpurun_out/${tag}_pytest.log; cut -c1-1800 gpurun_out/${tag}_bench_n1.log; tail -2 gpurun_out/${tag}_ncu.log; tail -5 gpurun_out/${tag}_memcheck.log
