#!/bin/bash
mkdir -p gpurun_out
for V in libstst_workloads_shfl.so ""; do echo "== ${V:-default}"; for W in jacobi5 hotspot; do STST_WORKLOADS_LIB=$V timeout 300 python scratch/sweep.py --workload $W --fuse 3,4,5 --iters 60 2>&1 | grep -v "^workload"; done; done > gpurun_out/sweep_shfl.log 2>&1; cat gpurun_out/sweep_shfl.log
