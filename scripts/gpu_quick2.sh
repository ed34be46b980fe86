#!/bin/bash
mkdir -p gpurun_out
for W in jacobi5 hotspot; do STST_WORKLOADS_LIB=libstst_workloads_t512.so timeout 300 python scratch/sweep.py --workload $W --fuse 4,6,8 --by 4,8 --ctas 2,3 --iters 48 2>&1 | grep -v "^workload"; done > gpurun_out/sweep_t512.log 2>&1; cat gpurun_out/sweep_t512.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sweep -s 4 -c 1 -f -o gpurun_out/prof_jacobi5_v3 python scratch/one.py --workload jacobi5 --iters 24 --calls 2 > gpurun_out/ncu_full_jacobi5_v3.log 2>&1
