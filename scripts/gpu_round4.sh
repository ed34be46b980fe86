#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --workload convection_pt --rows 8192 --cols 65536 --iterations 50 --steps 2 --warmup 1 > gpurun_out/bench_convection_big.json 2> gpurun_out/bench_convection_big.err; tail -5 gpurun_out/bench_convection_big.err; cat gpurun_out/bench_convection_big.json
