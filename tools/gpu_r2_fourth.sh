#!/bin/bash
# Round 2, fourth visit: new bench.py (regimes, parity, GSD e2e, sub-configs) + lane-group tracker queue.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_sizes.py tests/test_gpu_process_file.py tests/test_gpu_reference_tests.py -x -q > $OUT/r2d_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2d_pytest.log
tail -4 $OUT/r2d_pytest.log
timeout 1500 python bench.py > $OUT/r2d_bench.json 2> $OUT/r2d_bench.err; echo "bench exit $?"; head -c 6000 $OUT/r2d_bench.json; echo; tail -15 $OUT/r2d_bench.err
