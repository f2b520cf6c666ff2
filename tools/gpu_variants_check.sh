#!/bin/bash
# A/B of library variants built into var/ (every variant must print the same state checksum), then the GPU suite with
# the first variant, then memcheck of the list-overflow tests with the first variant.
# Usage on the GPU box: tools/gpu_variants_check.sh <tag> <variant> [variant ...]   -> gpurun_out/<tag>_*
tag=$1; shift
mkdir -p gpurun_out
for v in base "$@"; do
  SF_B200_LIB=$PWD/var/libsf_$v.so timeout 120 python tools/exp_bench.py 203 1500 40 2>&1 | tail -2
done | tee gpurun_out/${tag}_variants.log
SF_B200_LIB=$PWD/var/libsf_$1.so timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_$1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_$1.log
tail -3 gpurun_out/${tag}_pytest_$1.log
SF_B200_LIB=$PWD/var/libsf_$1.so timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "crowded or capacity or production_list" > gpurun_out/${tag}_memcheck_$1.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_memcheck_$1.log
tail -4 gpurun_out/${tag}_memcheck_$1.log
