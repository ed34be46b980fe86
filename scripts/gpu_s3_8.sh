#!/bin/bash
mkdir -p gpurun_out
{
echo "== fdtd, 512-thread CTAs (variant build)"
STST_WORKLOADS_LIB=libstst_workloads_mid512.so timeout 600 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 2,3,4 --iters 60 --ctas 1 --by 8,16 2>&1 | grep -v "^workload"
echo "== fdtd default"
timeout 600 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 0,3 --iters 60 --ctas 1 2>&1 | grep -v "^workload"
} > gpurun_out/sweep_fdtd512.log 2>&1; cat gpurun_out/sweep_fdtd512.log
for W in jacobi_r2 jacobi_r3 conway; do
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -3 gpurun_out/bench_$W.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$W.json').read().strip().splitlines()[-1]); print('$W', round(d['value'],1), round(d['roofline']['frac'],3), round(d['e2e']['value'],1), d['config']['fused_iterations'], d['config']['tile'], round(d['cpu_baseline']['value'],3))"
done
mkdir -p gpurun_out/driver2
timeout 900 python scripts/benchmark.py deep_grid_scaling hotspot --out-dir gpurun_out/driver2 --target-runtime 0.3 --max-wh 2048 2>&1 | tail -14
