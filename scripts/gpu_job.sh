mkdir -p gpurun_out/r02b
(time timeout 600 python scripts/fma_parity_probe.py) > gpurun_out/r02b/fma_probe.log 2>&1
for w in jacobi5 hotspot fdtd convection_pt jacobi_r3; do
  STST_STRICT=1 timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b/bench_strict_$w.json 2> gpurun_out/r02b/bench_strict_$w.err
done
(time timeout 1500 python -m pytest tests -m gpu -q -s --durations=15) > gpurun_out/r02b/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b/pytest_gpu.log
cat gpurun_out/r02b/fma_probe.log; tail -c 1500 gpurun_out/r02b/pytest_gpu.log
