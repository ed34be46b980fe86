#!/bin/bash
mkdir -p gpurun_out
for V in "" libstst_workloads_edge.so libstst_workloads_cw8.so libstst_workloads_cw8edge.so; do
  echo "=== variant: ${V:-default}"
  for W in jacobi5 hotspot; do
    STST_WORKLOADS_LIB=$V timeout 300 python scratch/sweep.py --workload $W --fuse 2,3,4,5 --iters 60 2>&1 | grep -v "^workload"
  done
done > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
