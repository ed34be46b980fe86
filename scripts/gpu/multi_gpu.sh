#!/bin/bash
# usage: bash scripts/gpu/multi_gpu.sh N   (run under `gpurun --gpus N`)
N=${1:-2}
out=gpurun_out/r02_n$N; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nvidia-smi topo -m > $out/topo.txt 2>&1; nproc >> $out/topo.txt; numactl -H >> $out/topo.txt 2>&1
if [ "$N" = "2" ]; then
  (time timeout 1200 python -m pytest tests/test_multi_device_update_gpu.py tests/test_sharding_gpu.py tests/test_apps_sharded.py tests/test_examples_gpu.py -m gpu -q) > $out/pytest_multi.log 2>&1
  echo "pytest rc=$?" >> $out/pytest_multi.log
fi
(time timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3) > $out/bench_default.json 2> $out/bench_default.err
DEV=$(python -c "print(','.join(str(i) for i in range($N)))")
(timeout 120 python scripts/run_one.py --workload hotspot --iters 1000 --calls 3 --devices $DEV) > $out/single_process_hotspot.log 2>&1
(timeout 120 python scripts/run_one.py --workload jacobi5 --iters 1002 --calls 3 --devices $DEV) > $out/single_process_jacobi5.log 2>&1
(timeout 120 python scripts/run_one.py --workload fdtd --rows 4608 --cols 4608 --iters 1000 --calls 3 --devices $DEV) > $out/single_process_fdtd.log 2>&1
(time timeout 200 $TR scripts/fdtd_max_grid.py) > $out/fdtd_max_grid.json 2> $out/fdtd_max_grid.err
(STST_HALO_TRANSPORT=nccl timeout 200 $TR bench.py --workload fdtd --scaling strong --steps 5 --warmup 3) > $out/bench_fdtd_nccl.json 2> $out/bench_fdtd_nccl.err
(timeout 100 $TR scripts/host_link_ceiling.py --gib 1) > $out/host_link_ceiling.json 2> $out/host_link_ceiling.err
(timeout 100 $TR scripts/host_link_ceiling.py --gib 1 --bind) > $out/host_link_ceiling_bound.json 2> $out/host_link_ceiling_bound.err
grep -E "passed|failed|rc=" $out/pytest_multi.log 2>/dev/null | tail -3; tail -c 600 $out/bench_default.err; cat $out/single_process_*.log; cat $out/host_link_ceiling*.json
