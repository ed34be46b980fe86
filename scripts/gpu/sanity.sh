#!/bin/bash
# half a minute on one GPU: smoke() and the small-grid / multi-slab parity tests on freshly built libraries
mkdir -p gpurun_out/r02_sanity
(python -c "import __graft_entry__ as g; g.smoke()" ; timeout 150 python -m pytest tests/test_parity_gpu.py tests/test_multi_device_update_gpu.py tests/test_sharding_gpu.py -m gpu -q -x -k "small_grid or sharded_call or (sharded_update and (hotspot or jacobi5 or fdtd) and not no-overlap and not 3-cuda)") > gpurun_out/r02_sanity/log.txt 2>&1
tail -4 gpurun_out/r02_sanity/log.txt
