#!/bin/bash
# round-2 GPU call 2 (re-entry): validate HEAD -- GPU tests, bench both arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2b_env.log; nproc >> gpurun_out/r2b_env.log; free -g | head -2 >> gpurun_out/r2b_env.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bench_n1.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bench_ref.log
tail -15 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_bench_n1.log | cut -c1-3000; cat gpurun_out/r2b_bench_ref.log | cut -c1-1500
