#!/bin/bash
mkdir -p gpurun_out
for v in coop; do
  SF_B200_LIB=$PWD/simplefluid_b200/lib/exp_$v.so timeout 300 python tools/exp_bench.py 203 1500 40 >> gpurun_out/r2i_exp.log 2>&1
done
cat gpurun_out/r2i_exp.log
