#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
{
for S in 1 0; do echo "== STST_SPECULATE=$S"
STST_SPECULATE=$S timeout 600 python scratch/sweep.py --workload hotspot --fuse 3,4,5,6 --ctas 2 --iters 60 2>&1 | grep -v "^workload"
STST_SPECULATE=$S timeout 600 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 2,3,4 --iters 60 --ctas 1 2>&1 | grep -v "^workload"
STST_SPECULATE=$S timeout 300 python scratch/sweep.py --workload convection_pt --rows 4096 --cols 8192 --fuse 1 --iters 10 --ctas 1 2>&1 | grep -v "^workload"
done
} > gpurun_out/sweep_spec.log 2>&1; cat gpurun_out/sweep_spec.log
