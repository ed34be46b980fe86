#!/bin/bash
# One GPU-box session: parity tests (incl. multi-process slabs), bench lines.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_jacobi5.json 2> gpurun_out/bench_jacobi5.err
cat gpurun_out/bench_jacobi5.json; tail -5 gpurun_out/bench_jacobi5.err
