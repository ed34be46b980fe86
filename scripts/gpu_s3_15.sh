#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "conway or smoke or degenerate or example or reference_cuda" > gpurun_out/pytest_conway.log 2>&1; tail -4 gpurun_out/pytest_conway.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --workload conway --steps 3 --warmup 3 > gpurun_out/bench_conway.json 2> gpurun_out/bench_conway.err; tail -2 gpurun_out/bench_conway.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_conway.json').read().strip().splitlines()[-1]); print('conway', round(d['value'],1), round(d['roofline']['frac'],3), round(d['e2e']['value'],1), d['config']['fused_iterations'], d['config']['tile'], d['config']['block'])"
