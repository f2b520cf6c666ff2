#!/bin/bash
# round-2 GPU call 12 (2 GPUs): multi-GPU parity suite incl. the Shepard slab cases, host-owned step, slab checkpoint
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --durations=10 > gpurun_out/r2l_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_2gpu.log
tail -25 gpurun_out/r2l_pytest_2gpu.log
