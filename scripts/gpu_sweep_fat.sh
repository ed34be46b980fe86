#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
timeout 300 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 0,1,2,3,4 --ctas 1,2 --iters 24 > gpurun_out/sweep_fdtd.log 2>&1; cat gpurun_out/sweep_fdtd.log
timeout 300 python scratch/sweep.py --workload convection_pt --rows 4096 --cols 8192 --fuse 0,1,2 --ctas 1,2 --iters 8 > gpurun_out/sweep_convection.log 2>&1; cat gpurun_out/sweep_convection.log
timeout 300 python scratch/sweep.py --workload hotspot --fuse 0,3,4 --ctas 1,2 --iters 48 > gpurun_out/sweep_hotspot2.log 2>&1; cat gpurun_out/sweep_hotspot2.log
timeout 300 python scratch/sweep.py --workload jacobi5 --fuse 0,3,4 --ctas 1,2 --iters 48 > gpurun_out/sweep_jacobi2.log 2>&1; cat gpurun_out/sweep_jacobi2.log
