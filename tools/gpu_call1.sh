#!/bin/bash
# round-2 GPU call 1: new parity tests, developed-flow bench, reference arm, ncu captures on the developed state
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2a_env.log; nproc >> gpurun_out/r2a_env.log; free -g | head -2 >> gpurun_out/r2a_env.log
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_bench_n1.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_bench_ref.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(density|force|visc)_brick' --launch-skip 4530 --launch-count 3 -f -o gpurun_out/r2a_dev_dam8m python tools/prof_run.py Dambreak 203 4 1510 > gpurun_out/r2a_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 18300 -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench_n1.log | cut -c1-1500
