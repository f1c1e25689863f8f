#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for T in 1 3 7 15; do SDB_COPY_THREADS=$T timeout 300 python tools/host_copy_probe.py > $OUT/r2g_copy_probe_$T.txt 2>&1; echo "probe $T exit $?"; grep "engine copy\|Register\|H2D\|numpy" $OUT/r2g_copy_probe_$T.txt; done
timeout 600 python -m pytest tests/test_gpu_cabi.py -x -q > $OUT/r2g_pytest_cabi.log 2>&1; echo "cabi exit $?"; tail -3 $OUT/r2g_pytest_cabi.log
