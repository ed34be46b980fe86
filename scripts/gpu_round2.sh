#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_examples_gpu.py -x -q > gpurun_out/pytest_examples.log 2>&1; tail -15 gpurun_out/pytest_examples.log
for W in fdtd convection_pt; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  tail -3 gpurun_out/bench_$W.err; cat gpurun_out/bench_$W.json
done
timeout 300 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 1,2,3,4 --iters 24 > gpurun_out/sweep_fdtd.log 2>&1; cat gpurun_out/sweep_fdtd.log
timeout 300 python scratch/sweep.py --workload convection_pt --rows 4096 --cols 8192 --fuse 1,2 --iters 8 > gpurun_out/sweep_convection.log 2>&1; cat gpurun_out/sweep_convection.log
# launch list of the bench command itself (contract): first 400 launches after warm-up
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 400 --csv --log-file gpurun_out/launches_bench_jacobi5.csv python bench.py --steps 1 --warmup 1 --iterations 300 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for W in hotspot fdtd convection_pt; do
  R=16384; C=16384; [ $W = fdtd ] && R=4608 && C=4608; [ $W = convection_pt ] && R=4096 && C=8192
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_sweep -s 4 -c 1 -f -o gpurun_out/prof_$W python scratch/one.py --workload $W --rows $R --cols $C --iters 12 --calls 2 > gpurun_out/ncu_full_$W.log 2>&1
done
ls -la gpurun_out
