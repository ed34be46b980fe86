#!/bin/bash
out=gpurun_out/r02_final; mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -q -s --durations=15) > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $out/clocks.csv &
SMI=$!
(time timeout 900 python bench.py --steps 10 --warmup 3) > $out/bench_default.json 2> $out/bench_default.err
for w in jacobi_r2 jacobi_r3 conway hotspot fdtd convection_pt; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 > $out/bench_$w.json 2> $out/bench_$w.err
done
kill $SMI
(time timeout 400 python bench.py --impl reference --steps 3 --warmup 1) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
bash scripts/gpu/ncu_captures.sh > $out/ncu_job.log 2>&1
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.log | tail -5; tail -c 400 $out/bench_default.err; tail -5 $out/ncu_job.log
