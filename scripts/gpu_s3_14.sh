#!/bin/bash
# final pass of session 3: whole GPU suite, bench lines of every workload, Conway column-group variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for W in jacobi5 hotspot fdtd convection_pt; do
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -3 gpurun_out/bench_$W.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$W.json').read().strip().splitlines()[-1]); print('$W', round(d['value'],1), round(d['roofline']['frac'],3), round(d['e2e']['value'],1), d['config']['fused_iterations'], d['config']['tile'], d['config']['passthrough_planes'], d['gpu_launches'])"
done
{
for V in "" libstst_workloads_b8.so libstst_workloads_b16.so; do echo "== conway ${V:-default}"
STST_WORKLOADS_LIB=$V timeout 300 python scratch/sweep.py --workload conway --fuse 0,8 --ctas 2 --iters 64 2>&1 | grep -v "^workload"
done
} > gpurun_out/sweep_conway.log 2>&1; cat gpurun_out/sweep_conway.log
