#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding_gpu.py tests/test_apps_sharded.py -m gpu -x -q -k "fdtd or hotspot or convection or passthrough" > gpurun_out/pytest_shard.log 2>&1; tail -4 gpurun_out/pytest_shard.log
