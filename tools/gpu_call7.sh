#!/bin/bash
mkdir -p gpurun_out
SF_B200_LIB=$PWD/simplefluid_b200/lib/exp_pool12.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_brick' --launch-skip 1510 --launch-count 1 -f -o gpurun_out/r2g_density_pool12 python tools/prof_run.py Dambreak 203 3 1510 > gpurun_out/r2g_ncu.log 2>&1
tail -3 gpurun_out/r2g_ncu.log
