#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_order.py -x -q > $OUT/r2w_pytest.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/r2w_pytest.log
