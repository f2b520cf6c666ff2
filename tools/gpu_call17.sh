#!/bin/bash
mkdir -p gpurun_out
SF_B200_LIB=$PWD/simplefluid_b200/lib/exp_waitstat.so timeout 300 python tools/exp_bench.py 203 1500 40 > gpurun_out/r2q_waitstat.log 2>&1
cat gpurun_out/r2q_waitstat.log
