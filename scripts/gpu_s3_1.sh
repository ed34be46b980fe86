#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/xfer.py > gpurun_out/xfer.log 2>&1; cat gpurun_out/xfer.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scatter|gather' -c 40 --csv --log-file gpurun_out/xfer_kernels.csv python scratch/xfer.py fdtd 4608 4608 > gpurun_out/xfer_ncu.log 2>&1
tail -15 gpurun_out/xfer_kernels.csv | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
