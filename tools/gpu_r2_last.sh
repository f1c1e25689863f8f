#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cabi.py tests/test_gpu_process_file.py tests/test_gpu_reference_tests.py tests/test_gpu_order.py -x -q > $OUT/r2y_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2y_pytest.log
timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -1
