#!/bin/bash
# ncu captures of the shipped sweep kernels (one GPU). Summaries -> gpurun_out/r02_ncu/, copied to profiles/.
out=gpurun_out/r02_ncu; mkdir -p $out
cap() { # name, skip, run_one args...
  name=$1; skip=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sweep_kernel -s $skip -c 1 -f -o $out/prof_$name \
      python scripts/run_one.py "$@" --calls 1 > $out/ncu_full_$name.log 2>&1
  python scripts/ncu_summary.py $out/prof_$name.ncu-rep > $out/ncu_${name}_summary.txt 2>&1
  ncu -i $out/prof_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $out/${name}_source_sass.csv.gz
  rm -f $out/prof_$name.ncu-rep
}
cap jacobi5 3 --workload jacobi5 --iters 48
cap hotspot 3 --workload hotspot --iters 32
cap fdtd 4 --workload fdtd --rows 4608 --cols 4608 --iters 32
cap convection_pt 2 --workload convection_pt --rows 4096 --cols 8192 --iters 6
cap jacobi_r3 3 --workload jacobi_r3 --iters 24
cap jacobi_r2 3 --workload jacobi_r2 --iters 32
# launch list of a bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_jacobi5.csv \
    python bench.py --workload jacobi5 --iterations 300 --steps 1 --warmup 1 --no-cpu-baseline > $out/launches_bench_jacobi5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_hotspot.csv \
    python bench.py --workload hotspot --iterations 200 --steps 1 --warmup 1 --no-cpu-baseline > $out/launches_bench_hotspot.log 2>&1
# one pass over the eight 576-row slabs of the FDTD max_grid grid, all on this GPU (what a slab pass consists of
# and what its launches cost: profiles/r02_launches_fdtd_8_slabs_one_gpu.csv)
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 64 -c 160 --csv --log-file $out/launches_fdtd_8_slabs_one_gpu.csv \
    python scripts/run_one.py --workload fdtd --rows 4608 --cols 4608 --iters 45 --calls 1 --devices 0,0,0,0,0,0,0,0 > $out/launches_fdtd_8_slabs_one_gpu.log 2>&1
# the reference drivers' figures through the scraper
for w in jacobi5 hotspot fdtd convection_pt; do
  mkdir -p $out/driver_$w
  if [ $w = convection_pt ]; then extra="--rows 4096 --cols 8192 --iterations 8"; else extra=""; fi
  timeout 600 python scripts/benchmark.py ncu_metrics $w --out-dir $out/driver_$w $extra > $out/driver_$w/scrape.log 2>&1
done
ls -la $out; head -8 $out/ncu_jacobi5_summary.txt; cat $out/driver_jacobi5/metrics.cuda.json
