#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "full_size_configs or reference_binarys" --durations=8 > gpurun_out/r2t_pytest_fullsize.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_fullsize.log
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench.log 2>&1
tail -14 gpurun_out/r2t_pytest_fullsize.log; grep -o '"parity_vs_reference_binary": [a-z]*' gpurun_out/r2t_bench.log; grep -o '"value": [0-9.e+]*' gpurun_out/r2t_bench.log | head -1
