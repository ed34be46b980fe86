mkdir -p gpurun_out/r02c
(time timeout 900 python -m pytest tests/test_parity_fullsize_gpu.py -q -s --durations=10) > gpurun_out/r02c/pytest_fullsize.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c/pytest_fullsize.log
for v in mid320 mid384; do
  STST_WORKLOADS_LIB=libstst_workloads_$v.so timeout 300 python bench.py --workload fdtd --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c/bench_fdtd_$v.json 2> gpurun_out/r02c/bench_fdtd_$v.err
done
(time timeout 600 python scripts/fdtd_max_grid.py) > gpurun_out/r02c/fdtd_max_grid_1gpu.json 2> gpurun_out/r02c/fdtd_max_grid_1gpu.err
(time bash scripts/sanitize.sh) > gpurun_out/r02c/sanitize.log 2>&1
grep -E "passed|failed|rc=|fullsize parity" gpurun_out/r02c/pytest_fullsize.log | tail -20; cat gpurun_out/r02c/sanitize.log | tail -30
