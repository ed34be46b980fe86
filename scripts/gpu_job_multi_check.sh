#!/bin/bash
N=${1:-2}
out=gpurun_out/r02_n${N}_check; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
(time timeout 600 python -m pytest tests/test_multi_device_update_gpu.py tests/test_sharding_gpu.py tests/test_reference_unit_tests_gpu.py -m gpu -q -x -k "sharded_call or environment or small_grids or nccl or update_extensions") > $out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $out/pytest_multi.log
DEV=$(python -c "print(','.join(str(i) for i in range($N)))")
(timeout 200 python scripts/run_one.py --workload hotspot --iters 1000 --calls 3 --devices $DEV) > $out/single_process_hotspot.log 2>&1
(timeout 200 python scripts/run_one.py --workload jacobi5 --iters 1002 --calls 3 --devices $DEV) > $out/single_process_jacobi5.log 2>&1
(timeout 200 python scripts/run_one.py --workload fdtd --rows 4608 --cols 4608 --iters 1000 --calls 3 --devices $DEV) > $out/single_process_fdtd.log 2>&1
(STST_HALO_TRANSPORT=nccl timeout 200 $TR bench.py --workload fdtd --scaling strong --steps 5 --warmup 3) > $out/bench_fdtd_nccl.json 2> $out/bench_fdtd_nccl.err
(STST_HALO_TRANSPORT=nccl timeout 200 $TR bench.py --workload hotspot --scaling strong --steps 3 --warmup 3) > $out/bench_hotspot_nccl.json 2> $out/bench_hotspot_nccl.err
grep -E "passed|failed|rc=" $out/pytest_multi.log | tail -3; cat $out/single_process_*.log | tail -12; tail -c 300 $out/bench_fdtd_nccl.json
