#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811"
SF_SLAB_TRACE=2 timeout 900 $TR bench.py --gpus 8 --steps 100 --settle 300 --no-verify > gpurun_out/r2p_bench_n8_timeline.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_bench_n8_timeline.log
grep "device timeline" gpurun_out/r2p_bench_n8_timeline.log | sort | cut -c1-700; grep -o '"ms_per_step[a-z_]*": [0-9.]*' gpurun_out/r2p_bench_n8_timeline.log
