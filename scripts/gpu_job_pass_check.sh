#!/bin/bash
out=gpurun_out/r02_pass1; mkdir -p $out
(time timeout 900 python -m pytest tests/test_sharding_gpu.py tests/test_multi_device_update_gpu.py tests/test_apps_sharded.py tests/test_examples_gpu.py tests/test_reference_unit_tests_gpu.py -m gpu -q) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
(STST_SLAB_PASS=split timeout 600 python -m pytest tests/test_sharding_gpu.py tests/test_multi_device_update_gpu.py -m gpu -q -k "hotspot or fdtd or jacobi5") > $out/pytest_split.log 2>&1
echo "pytest rc=$?" >> $out/pytest_split.log
grep -E "passed|failed|FAILED|rc=" $out/pytest.log $out/pytest_split.log | tail -8
