#!/bin/bash
# A/B timing of library variants built into var/ (development aid): every variant must print the same state hash
mkdir -p gpurun_out
for v in "$@"; do
  SF_B200_LIB=$PWD/var/libsf_$v.so timeout 300 python tools/exp_bench.py 203 1500 40 2>&1 | tail -4
done | tee gpurun_out/exp_variants.log
