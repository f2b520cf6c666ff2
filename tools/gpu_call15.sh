#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29911"
run() { name=$1; shift; env "$@" SF_SLAB_TRACE=1 timeout 600 $TR bench.py --gpus 4 --steps 60 --settle 100 --no-verify > gpurun_out/r2o_$name.log 2>&1; echo "== $name: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2o_$name.log) $(grep -o '"ms_per_step_at_rest": [0-9.]*' gpurun_out/r2o_$name.log)"; grep "slab rank 1\]" gpurun_out/r2o_$name.log | cut -c1-260; }
run base NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P
run chan2 NCCL_MAX_P2P_NCHANNELS=2
run free16 SF_SLAB_FREE_SLOTS=16
run free16chan8 SF_SLAB_FREE_SLOTS=16 NCCL_MAX_P2P_NCHANNELS=8
grep -iE "via P2P|NVLS|channels|P2P Chunk|nChannels" gpurun_out/r2o_base.log | head -12
