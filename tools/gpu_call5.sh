#!/bin/bash
# round-2 GPU call 5: kernel experiments (wait back-off, table via L1, small bricks)
mkdir -p gpurun_out
for v in base backoff tabglobal bz2 bz2backoff; do
  SF_B200_LIB=$PWD/simplefluid_b200/lib/exp_$v.so timeout 300 python tools/exp_bench.py 203 1500 40 >> gpurun_out/r2e_exp.log 2>&1
done
cat gpurun_out/r2e_exp.log
