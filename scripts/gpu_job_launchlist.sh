#!/bin/bash
out=gpurun_out/r02_last; mkdir -p $out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 64 -c 160 --csv --log-file $out/launches_fdtd_8_slabs_one_gpu.csv \
    python scripts/run_one.py --workload fdtd --rows 4608 --cols 4608 --iters 45 --calls 1 --devices 0,0,0,0,0,0,0,0 > $out/launches_fdtd_8_slabs_one_gpu.log 2>&1
tail -3 $out/launches_fdtd_8_slabs_one_gpu.log; wc -l $out/launches_fdtd_8_slabs_one_gpu.csv
