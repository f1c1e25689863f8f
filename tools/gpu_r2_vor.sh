#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_voronoi.py tests/test_gpu_process_file.py -m gpu -x -q > $OUT/r2v_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2v_pytest.log
timeout 120 python tools/aux_bench.py 2>&1 | grep voronoi
