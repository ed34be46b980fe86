#!/bin/bash
mkdir -p gpurun_out
for S in 1 0; do
STST_SPECULATE=$S timeout 600 python bench.py --workload fdtd --steps 3 --warmup 3 --chunk-above-gib 0 --no-cpu-baseline > gpurun_out/bench_fdtd_slab_$S.json 2> gpurun_out/bench_fdtd_slab_$S.err; tail -3 gpurun_out/bench_fdtd_slab_$S.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_fdtd_slab_$S.json').read().strip().splitlines()[-1]); print('fdtd slab path, speculate=$S:', round(d['value'],1), d['ms_per_step'], d['config'])"
done
STST_TRACE=1 timeout 300 python scratch/sweep.py --workload fdtd --rows 4608 --cols 4608 --fuse 2,3 --iters 60 --ctas 2,1 2>&1 | grep -v "^workload\|stst" | tail -6
