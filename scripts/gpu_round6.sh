#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for W in jacobi5 hotspot fdtd convection_pt; do
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -3 gpurun_out/bench_$W.err; cut -c1-330 gpurun_out/bench_$W.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sweep -s 4 -c 1 -f -o gpurun_out/prof_jacobi5_v4 python scratch/one.py --workload jacobi5 --iters 24 --calls 2 > gpurun_out/ncu_full_jacobi5_v4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file gpurun_out/launches_bench_jacobi5.csv python bench.py --steps 1 --warmup 1 --iterations 300 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
