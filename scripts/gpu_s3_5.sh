#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python scratch/sweep.py --workload jacobi5 --fuse 5,6,7 --ctas 2,3 --iters 120 2>&1 | grep -v "^workload"
timeout 600 python scratch/sweep.py --workload hotspot --fuse 3,4,5 --ctas 2 --iters 60 2>&1 | grep -v "^workload"
timeout 600 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 2,3 --iters 60 --ctas 1 2>&1 | grep -v "^workload"
timeout 300 python scratch/sweep.py --workload convection_pt --rows 4096 --cols 8192 --fuse 1 --iters 10 --ctas 1 2>&1 | grep -v "^workload"
} > gpurun_out/sweep_s3_5.log 2>&1; cat gpurun_out/sweep_s3_5.log
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
