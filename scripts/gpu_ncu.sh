#!/bin/bash
# bash scripts/gpu_ncu.sh "workload rows cols iters [fuse]" ...
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec
  W=$1; R=$2; C=$3; I=$4; F=${5:-0}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_sweep -s 4 -c 1 -f -o gpurun_out/prof_$W python scratch/one.py --workload $W --rows $R --cols $C --iters $I --fuse $F --calls 2 > gpurun_out/ncu_full_$W.log 2>&1
  tail -2 gpurun_out/ncu_full_$W.log
done
