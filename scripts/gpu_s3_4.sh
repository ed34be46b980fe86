#!/bin/bash
mkdir -p gpurun_out
{
echo "== cw8 variant (lane-major, 32 threads x 8 columns)"
STST_WORKLOADS_LIB=libstst_workloads_cw8.so timeout 600 python scratch/sweep.py --workload jacobi5 --fuse 4,5,6,8 --ctas 2,3 --iters 120 2>&1 | grep -v "^workload"
echo "== default"
timeout 600 python scratch/sweep.py --workload jacobi5 --fuse 6 --ctas 2 --iters 120 2>&1 | grep -v "^workload"
} > gpurun_out/sweep_cw8.log 2>&1; cat gpurun_out/sweep_cw8.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_sweep -s 4 -c 1 -f -o /tmp/prof_jacobi5_s3 python scratch/one.py --workload jacobi5 --iters 24 --calls 2 > gpurun_out/ncu_full_jacobi5_s3.log 2>&1
tail -2 gpurun_out/ncu_full_jacobi5_s3.log
ls -la /tmp/prof_jacobi5_s3.ncu-rep
ncu -i /tmp/prof_jacobi5_s3.ncu-rep --page source --csv --print-source sass > gpurun_out/jacobi5_s3_source_sass.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/prof_jacobi5_s3.ncu-rep > gpurun_out/jacobi5_s3_summary.txt 2>&1
gzip -f gpurun_out/jacobi5_s3_source_sass.csv; ls -la gpurun_out/
[ $(stat -c %s /tmp/prof_jacobi5_s3.ncu-rep) -lt 30000000 ] && cp /tmp/prof_jacobi5_s3.ncu-rep gpurun_out/
