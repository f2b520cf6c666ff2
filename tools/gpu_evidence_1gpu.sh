#!/bin/bash
# round-2 GPU call 11: final single-GPU evidence -- GPU suite, bench both arms, ncu full capture (developed state), launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
timeout 900 python bench.py > gpurun_out/r2x_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_bench_n1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(density|force|visc)_brick' --launch-skip 4530 --launch-count 3 -f -o gpurun_out/r2x_dev_dam8m python tools/prof_run.py Dambreak 203 4 1510 > gpurun_out/r2x_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 18300 -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_bench_under_ncu.log 2>&1
tail -4 gpurun_out/r2x_pytest.log; cut -c1-1800 gpurun_out/r2x_bench_n1.log; tail -2 gpurun_out/r2x_ncu.log
