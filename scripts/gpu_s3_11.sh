#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding_gpu.py -x -q > gpurun_out/pytest_multi.log 2>&1; tail -5 gpurun_out/pytest_multi.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for W in hotspot fdtd; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload $W --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale_${W}_2.json 2> gpurun_out/scale_${W}_2.err
  tail -3 gpurun_out/scale_${W}_2.err | cut -c1-300
  python -c "
import json; d=json.loads(open('gpurun_out/scale_${W}_2.json').read().strip().splitlines()[-1]); print('$W N=2', round(d['value'],1), d['ms_per_step'], round(d['e2e']['value'],1), d['e2e'].get('host_memory'), d['config']['tile'], d['gpu_launches'])"
done
