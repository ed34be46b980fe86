#!/bin/bash
N=${1:-8}
out=gpurun_out/r02_pass_ab_n$N; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for w in fdtd hotspot; do
  (timeout 200 $TR bench.py --workload $w --scaling strong --steps 5 --warmup 3) > $out/bench_${w}_single.json 2> $out/bench_${w}_single.err
  (STST_SLAB_PASS=split timeout 200 $TR bench.py --workload $w --scaling strong --steps 5 --warmup 3) > $out/bench_${w}_split.json 2> $out/bench_${w}_split.err
done
DEV=$(python -c "print(','.join(str(i) for i in range($N)))")
(timeout 120 python scripts/run_one.py --workload fdtd --rows 4608 --cols 4608 --iters 1000 --calls 3 --devices $DEV) > $out/single_process_fdtd.log 2>&1
(timeout 120 python scripts/run_one.py --workload hotspot --iters 1000 --calls 3 --devices $DEV) > $out/single_process_hotspot.log 2>&1
for f in $out/bench_*.json; do python -c "
import json,sys
r=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(r['value'],1), 'e2e', round(r['e2e']['value'],1), r['gpu_launches'], (r.get('parity') or {}).get('rel_max_norm'))"; done; tail -2 $out/single_process_*.log
