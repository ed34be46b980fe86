#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_field_ops_gpu.py tests/test_apps_gpu.py tests/test_sharding_gpu.py -x -q > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log
timeout 300 python scratch/reduce_bench.py 2>&1 | tail -8
for W in hotspot jacobi5; do
timeout 600 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -3 gpurun_out/bench_$W.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$W.json').read().strip().splitlines()[-1]); print('$W', d['value'], d['e2e'])"
done
STST_PIN_LIMIT_MB=64 timeout 600 python bench.py --workload jacobi5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_jacobi5_staged.json 2> gpurun_out/bench_jacobi5_staged.err; tail -3 gpurun_out/bench_jacobi5_staged.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_jacobi5_staged.json').read().strip().splitlines()[-1]); print('staged', d['value'], d['e2e'])"
STST_PIN_LIMIT_MB=64 timeout 300 python scratch/xfer.py hotspot 16384 16384 2>&1 | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
