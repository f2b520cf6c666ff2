#!/bin/bash
# round-2 GPU call 13 (8 GPUs): weak-scaling bench (64 M dambreak) with per-rank slab trace, strong-scaling bench (configs[2]), 4-rank parity case
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.log 2>&1; nproc >> gpurun_out/r2m_topo.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811"
SF_SLAB_TRACE=1 timeout 1200 $TR bench.py --gpus 8 > gpurun_out/r2m_bench_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_bench_n8.log
SF_SLAB_TRACE=1 timeout 900 $TR bench.py --gpus 8 --workload doubledambreak_8m > gpurun_out/r2m_bench_strong_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_bench_strong_n8.log
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -k "4-DoubleDambreak" > gpurun_out/r2m_pytest_4gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest_4gpu.log
for f in r2m_bench_n8 r2m_bench_strong_n8; do echo "== $f"; grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM" gpurun_out/$f.log | cut -c1-1900 | tail -12; done; tail -3 gpurun_out/r2m_pytest_4gpu.log
