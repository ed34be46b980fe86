#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_field_ops_gpu.py tests/test_apps_gpu.py tests/test_sharding_gpu.py tests/test_examples_gpu.py -x -q > gpurun_out/pytest_new.log 2>&1; tail -5 gpurun_out/pytest_new.log
timeout 300 python scratch/reduce_bench.py 2>&1 | tail -8
{
for V in "" libstst_workloads_r2.so libstst_workloads_r3.so; do echo "== ${V:-default}";
 STST_WORKLOADS_LIB=$V timeout 300 python scratch/sweep.py --workload convection_pt --rows 4096 --cols 8192 --fuse 1 --iters 10 --ctas 1 2>&1 | grep -v "^workload"
 STST_WORKLOADS_LIB=$V timeout 300 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 2,3 --iters 60 --ctas 1,2 2>&1 | grep -v "^workload"
done
echo "== ctas sweep"
timeout 600 python scratch/sweep.py --workload jacobi5 --fuse 3,4,5,6 --ctas 2,3,4 --iters 120 2>&1 | grep -v "^workload"
timeout 600 python scratch/sweep.py --workload hotspot --fuse 2,3,4 --ctas 2,3,4 --iters 60 2>&1 | grep -v "^workload"
} > gpurun_out/sweep_s3.log 2>&1; cat gpurun_out/sweep_s3.log
