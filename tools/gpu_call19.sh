#!/bin/bash
mkdir -p gpurun_out
for v in libsf_b200 exp_ahead8; do
  SF_B200_LIB=$PWD/simplefluid_b200/lib/$v.so timeout 300 python tools/exp_bench.py 203 1500 40 >> gpurun_out/r2s_exp.log 2>&1
done
cat gpurun_out/r2s_exp.log
