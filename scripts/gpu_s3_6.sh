#!/bin/bash
# round-1 session-3 measurement pass: bench lines, ncu launch list + full capture, benchmark driver
mkdir -p gpurun_out
for W in jacobi5 hotspot fdtd convection_pt; do
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -3 gpurun_out/bench_$W.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$W.json').read().strip().splitlines()[-1]); print('$W', round(d['value'],1), round(d['roofline']['frac'],3), round(d['e2e']['value'],1), d['e2e']['host_memory'], round(d['cpu_baseline']['value'],3), d['clocks'])"
done
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>&1; cut -c1-300 gpurun_out/bench_reference.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sweep -s 4 -c 1 -f -o gpurun_out/prof_jacobi5_v5 python scratch/one.py --workload jacobi5 --iters 24 --calls 2 > gpurun_out/ncu_full_jacobi5_v5.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_jacobi5_v5.ncu-rep > gpurun_out/ncu_jacobi5_v5_summary.txt 2>&1; head -8 gpurun_out/ncu_jacobi5_v5_summary.txt
for spec in "hotspot 16384 16384 16" "fdtd 4608 4608 12" "convection_pt 4096 8192 4"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sweep -s 4 -c 1 -f -o /tmp/prof_$1 python scratch/one.py --workload $1 --rows $2 --cols $3 --iters $4 --calls 2 > gpurun_out/ncu_full_$1_v5.log 2>&1
  python scripts/ncu_summary.py /tmp/prof_$1.ncu-rep > gpurun_out/ncu_$1_v5_summary.txt 2>&1; head -4 gpurun_out/ncu_$1_v5_summary.txt
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$1_v5_source_sass.csv.gz
done
timeout 300 ncu --set full --clock-control none -k regex:reduce_max_abs -c 1 -f -o /tmp/prof_reduce python scratch/reduce_bench.py > gpurun_out/ncu_full_reduce.log 2>&1
python scripts/ncu_summary.py /tmp/prof_reduce.ncu-rep > gpurun_out/ncu_reduce_max_abs_summary.txt 2>&1; head -6 gpurun_out/ncu_reduce_max_abs_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file gpurun_out/launches_bench_jacobi5_v5.csv python bench.py --steps 1 --warmup 1 --iterations 300 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -4 gpurun_out/launches_bench_jacobi5_v5.csv | cut -c1-260
mkdir -p gpurun_out/driver
timeout 600 python scripts/benchmark.py max_perf hotspot --out-dir gpurun_out/driver 2>&1 | tail -2
timeout 900 python scripts/benchmark.py deep_grid_scaling hotspot --out-dir gpurun_out/driver --target-runtime 0.4 --max-wh 23171 2>&1 | tail -25
