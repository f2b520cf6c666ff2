#!/bin/bash
# 2-GPU run: multi-GPU parity suite (slab runs, Shepard with four ghost layers, host-owned step, slab checkpoint) and the
# weak-scaling bench at N = 2 (prints parity_vs_single_gpu).  Usage: tools/gpu_tests_2gpu.sh [tag] -> gpurun_out/<tag>_*
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --durations=10 > gpurun_out/${tag}_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_2gpu.log
tail -8 gpurun_out/${tag}_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/${tag}_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_bench_n2.log
grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${tag}_bench_n2.log | cut -c1-1500 | tail -4
