#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --workload convection_pt --rows 8192 --cols 65536 --iterations 50 --steps 2 --warmup 1 > gpurun_out/bench_convection_big.json 2> gpurun_out/bench_convection_big.err; tail -5 gpurun_out/bench_convection_big.err; cat gpurun_out/bench_convection_big.json
for W in convection_pt fdtd hotspot jacobi5; do
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -3 gpurun_out/bench_$W.err; cat gpurun_out/bench_$W.json
done
