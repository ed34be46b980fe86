mkdir -p gpurun_out/r02d
for v in fat640 fat768; do
  STST_WORKLOADS_LIB=libstst_workloads_$v.so timeout 300 python bench.py --workload convection_pt --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02d/bench_convection_$v.json 2> gpurun_out/r02d/bench_convection_$v.err
done
(STST_WORKLOADS_LIB=libstst_workloads_fat640.so timeout 300 python scripts/sweep_plans.py --workload convection_pt --rows 4096 --cols 8192 --iters 20 --fuse 1 --ctas 2 --bx 32 --by 10) > gpurun_out/r02d/sweep_convection_fat640_2cta.log 2>&1
(STST_WORKLOADS_LIB=libstst_workloads_fat768.so timeout 300 python scripts/sweep_plans.py --workload convection_pt --rows 4096 --cols 8192 --iters 20 --fuse 1 --ctas 2 --bx 32 --by 12) > gpurun_out/r02d/sweep_convection_fat768_2cta.log 2>&1
timeout 300 python bench.py --workload convection_pt --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02d/bench_convection_base.json 2> gpurun_out/r02d/bench_convection_base.err
timeout 300 python bench.py --workload fdtd --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02d/bench_fdtd.json 2> gpurun_out/r02d/bench_fdtd.err
timeout 300 python bench.py --workload hotspot --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02d/bench_hotspot.json 2> gpurun_out/r02d/bench_hotspot.err
(timeout 600 python scripts/sweep_plans.py --workload jacobi_r3 --iters 48 --fuse 2,3,4 --ctas 1,2,3) > gpurun_out/r02d/sweep_jacobi_r3.log 2>&1
(timeout 600 python scripts/sweep_plans.py --workload jacobi_r2 --iters 48 --fuse 3,4,5 --ctas 2,3) > gpurun_out/r02d/sweep_jacobi_r2.log 2>&1
(time timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_reference_unit_tests_gpu.py tests/test_sharding_gpu.py -q -x -k "fdtd or kat or reference or own") > gpurun_out/r02d/pytest_subset.log 2>&1
tail -3 gpurun_out/r02d/pytest_subset.log
