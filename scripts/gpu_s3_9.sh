#!/bin/bash
mkdir -p gpurun_out
{
for G in 1024 512 2048; do
for W in hotspot jacobi5; do
timeout 300 python scratch/sweep.py --workload $W --rows $G --cols $G --fuse 2,4 --ctas 2,4,6 --tile_rows 0,12,6 --iters 4000 2>&1 | grep -v "^workload"
done; done
} > gpurun_out/sweep_small.log 2>&1; cat gpurun_out/sweep_small.log
