#!/bin/bash
mkdir -p gpurun_out
for W in fdtd hotspot; do
timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --chunk-above-gib 0 --no-cpu-baseline > gpurun_out/bench_${W}_slab.json 2> gpurun_out/bench_${W}_slab.err; tail -3 gpurun_out/bench_${W}_slab.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_${W}_slab.json').read().strip().splitlines()[-1]); print('$W slab path:', round(d['value'],1), d['ms_per_step'], d['config']['tile'], d['config']['passthrough_planes'], d['config']['speculation_redos'], round(d['e2e']['value'],1))"
done
timeout 300 python scratch/sweep.py --workload convection_pt --rows 4096 --cols 8192 --fuse 1 --iters 10 --ctas 1 2>&1 | grep -v "^workload"
timeout 300 python scratch/sweep.py --workload jacobi5 --fuse 6 --ctas 2 --iters 120 2>&1 | grep -v "^workload"
