#!/bin/bash
# compute-sanitizer over the tile kernels (memcheck + racecheck on the smem-tile
# kernels). Run on the GPU box; logs go to gpurun_out/sanitizer/, summaries are copied to profiles/.
# What it targets: the unclamped edge-column loads that rely on tile_guard_bytes, the per-plane buffer
# aliasing of the pass-through kernels, the lane-major conversion sweep, the non-TMA staging path on
# slabs (rows beyond the mapped planes), and the in-process multi-slab path with its peer stores.
out=gpurun_out/sanitizer; mkdir -p $out
SAN=/usr/local/cuda/bin/compute-sanitizer
sel_small='small_grid_matches_oracle or degenerate_and_odd_shapes or passthrough_planes_are_detected or wrong_speculation'
run() { # name, tool args..., -- pytest args...
  name=$1; shift
  ( time timeout 1500 $SAN "$@" ) > $out/$name.log 2>&1
  echo "exit code $?" >> $out/$name.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" $out/$name.log | tail -5
}
run memcheck_parity --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -q -x -k "$sel_small"
STST_TMA=0 run memcheck_parity_no_tma --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -q -x -k "$sel_small"
run memcheck_multi_slab --tool memcheck --error-exitcode 7 python -m pytest tests/test_multi_device_update_gpu.py -q -x -k "sharded_call"
run racecheck_parity --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -q -x -k "small_grid_matches_oracle"
STST_TMA=0 run racecheck_parity_no_tma --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -q -x -k "small_grid_matches_oracle"
run synccheck_parity --tool synccheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -q -x -k "small_grid_matches_oracle"
