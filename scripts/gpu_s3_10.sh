#!/bin/bash
mkdir -p gpurun_out gpurun_out/driver3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for W in hotspot fdtd jacobi5; do
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -3 gpurun_out/bench_$W.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$W.json').read().strip().splitlines()[-1]); print('$W', round(d['value'],1), round(d['roofline']['frac'],3), round(d['e2e']['value'],1), d['config']['fused_iterations'], d['config']['tile'], d['gpu_launches'])"
done
timeout 900 python scripts/benchmark.py deep_grid_scaling hotspot --out-dir gpurun_out/driver3 --target-runtime 0.3 --max-wh 23171 2>&1 | tail -21
for spec in "hotspot 16384 16384 16" "fdtd 4608 4608 12"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none -k regex:fused_sweep -s 6 -c 1 -f -o /tmp/prof_$1 python scratch/one.py --workload $1 --rows $2 --cols $3 --iters $4 --calls 3 > gpurun_out/ncu_full_$1_v6.log 2>&1
  python scripts/ncu_summary.py /tmp/prof_$1.ncu-rep > gpurun_out/ncu_$1_v6_summary.txt 2>&1; head -14 gpurun_out/ncu_$1_v6_summary.txt
done
