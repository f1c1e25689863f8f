#!/bin/bash
# Short single-GPU visit: the quick parity files of the suite, smoke, and the configs[0] sub-record on its own.
# usage: bash tools/gpu_check.sh [config0-only]
OUT=gpurun_out; mkdir -p $OUT
if [ "$1" != "config0-only" ]; then
  timeout 150 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_process_file.py tests/test_gpu_parity.py tests/test_gpu_reference_tests.py -x -q > $OUT/r2x_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2x_pytest.log
  timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -1
fi
SDB_BENCH_CONFIG0_ROWS=$OUT/r2x_config0_rows.csv timeout 120 python bench_configs.py 0_small > $OUT/r2x_config0.json 2> $OUT/r2x_config0.err; echo "config0 exit $?"; tail -c 900 $OUT/r2x_config0.json; tail -3 $OUT/r2x_config0.err
