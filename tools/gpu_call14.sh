#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29911"
SF_SLAB_TRACE=1 timeout 900 $TR bench.py --gpus 4 --steps 60 --settle 300 --no-verify > gpurun_out/r2n_bench_n4.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_bench_n4.log
grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2n_bench_n4.log | cut -c1-700 | tail -8
