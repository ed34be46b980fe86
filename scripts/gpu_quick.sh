#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
for W in jacobi5 hotspot; do timeout 300 python scratch/sweep.py --workload $W --fuse 0,3,4,5,6 --iters 60 2>&1 | grep -v "^workload"; done > gpurun_out/sweep_light.log 2>&1; cat gpurun_out/sweep_light.log
timeout 300 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 0,2,3 --ctas 1,2 --iters 24 2>&1 | grep -v "^workload" > gpurun_out/sweep_fdtd.log; cat gpurun_out/sweep_fdtd.log
timeout 300 python scratch/sweep.py --workload convection_pt --rows 4096 --cols 8192 --fuse 0 --iters 8 2>&1 | grep -v "^workload" > gpurun_out/sweep_convection.log; cat gpurun_out/sweep_convection.log
for W in conway jacobi_r3; do :; done
