#!/bin/bash
# Multi-GPU session: bash scripts/gpu_multi.sh N "workloads" [pytest]
set -x
mkdir -p gpurun_out
N=${1:-2}
WL=${2:-"jacobi5 hotspot"}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_multi.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
if [ "$3" = "pytest" ]; then
  timeout 900 python -m pytest tests/test_sharding_gpu.py -x -q -k "2-cuda" > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
fi
for W in $WL; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload $W --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_${W}_$N.json 2> gpurun_out/scale_${W}_$N.err
  tail -3 gpurun_out/scale_${W}_$N.err | cut -c1-300
  cat gpurun_out/scale_${W}_$N.json
  [ "$4" = "noverlap" ] && timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --workload $W --gpus $N --steps 3 --warmup 3 --no-overlap > gpurun_out/scale_${W}_${N}_noverlap.json 2> gpurun_out/scale_${W}_${N}_noverlap.err
  cut -c1-200 gpurun_out/scale_${W}_${N}_noverlap.json
done
