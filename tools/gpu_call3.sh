#!/bin/bash
# round-2 GPU call 3 (2 GPUs): multi-GPU parity incl. host-owned step and slab checkpoint; weak + strong bench at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2c_env.log; nproc >> gpurun_out/r2c_env.log; nvidia-smi topo -m >> gpurun_out/r2c_env.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
SF_SLAB_TRACE=1 timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2c_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_bench_n2.log
SF_SLAB_TRACE=1 timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 --workload doubledambreak_8m > gpurun_out/r2c_bench_strong_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_bench_strong_n2.log
timeout 600 python bench.py --steps 50 --warmup 5 --workload doubledambreak_8m --no-cpu-baseline > gpurun_out/r2c_bench_strong_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_bench_strong_n1.log
tail -12 gpurun_out/r2c_pytest.log; for f in r2c_bench_n2 r2c_bench_strong_n2 r2c_bench_strong_n1; do echo "== $f"; grep -v "^\[W\|^W0\|^\*\*\*" gpurun_out/$f.log | cut -c1-2500 | tail -8; done
