#!/bin/bash
# 8-GPU run: weak-scaling bench (64 M dambreak) with per-rank exchange statistics (SF_SLAB_TRACE=1), optionally the
# strong-scaling bench (configs[2]).  Usage: tools/gpu_bench_8gpu.sh [tag] [strong]  -> gpurun_out/<tag>_*
tag=${1:-r2d}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.log 2>&1; nproc >> gpurun_out/${tag}_topo.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811"
SF_SLAB_TRACE=1 timeout 1200 $TR bench.py --gpus 8 --no-cpu-baseline > gpurun_out/${tag}_bench_n8.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_bench_n8.log
if [ "$2" = "strong" ]; then
  SF_SLAB_TRACE=1 timeout 900 $TR bench.py --gpus 8 --no-cpu-baseline --workload doubledambreak_8m > gpurun_out/${tag}_bench_strong_n8.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_bench_strong_n8.log
fi
for f in ${tag}_bench_n8 ${tag}_bench_strong_n8; do if [ -f gpurun_out/$f.log ]; then echo "== $f"; grep -v "^\[W\|^W0\|OMP_NUM" gpurun_out/$f.log | cut -c1-1900 | tail -14; fi; done; exit 0
