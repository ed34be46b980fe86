mkdir -p gpurun_out/r02_sanity
(python -c "import __graft_entry__ as g; g.smoke()" ; timeout 100 python -m pytest tests/test_parity_gpu.py tests/test_multi_device_update_gpu.py -m gpu -q -x -k "small_grid or sharded_call_equals") > gpurun_out/r02_sanity/log.txt 2>&1
tail -4 gpurun_out/r02_sanity/log.txt
