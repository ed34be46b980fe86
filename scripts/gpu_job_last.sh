#!/bin/bash
out=gpurun_out/r02_last; mkdir -p $out
(time timeout 560 python -m pytest tests -m gpu -q -x) > $out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 60 python bench.py --workload jacobi_r2 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_jacobi_r2.json 2> $out/bench_jacobi_r2.err
timeout 60 python bench.py --workload conway --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_conway.json 2> $out/bench_conway.err
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.log | tail -4
