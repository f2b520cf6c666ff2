#!/bin/bash
# round-2 GPU call 6: pooled exact phase of the density pass -- variants + parity
mkdir -p gpurun_out
for v in pool0 pool4s pool9 pool12; do
  SF_B200_LIB=$PWD/simplefluid_b200/lib/exp_$v.so timeout 300 python tools/exp_bench.py 203 1500 40 >> gpurun_out/r2f_exp.log 2>&1
done
cat gpurun_out/r2f_exp.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -5 gpurun_out/r2f_pytest.log
