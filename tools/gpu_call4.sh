#!/bin/bash
# round-2 GPU call 4: GPU suite incl. the tests against the reference binary's outputs, smoke, source-level ncu capture of the density pass (developed state)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_density_brick' --launch-skip 1510 --launch-count 1 -f -o gpurun_out/r2d_density_dev python tools/prof_run.py Dambreak 203 3 1510 > gpurun_out/r2d_ncu.log 2>&1
tail -6 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_smoke.log; tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/*.ncu-rep
