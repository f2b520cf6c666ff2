#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_bench.py 203 1500 40 > gpurun_out/r2r_exp.log 2>&1
cat gpurun_out/r2r_exp.log
