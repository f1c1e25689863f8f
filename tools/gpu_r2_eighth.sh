#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for PIN in 0 1; do
SDB_BENCH_HOSTPROF=1 SDB_BENCH_PIN_FILE=$PIN timeout 600 python bench.py --regime aged --no-configs --no-cpu-baseline --min-seconds 0.4 > $OUT/r2h_bench_pin$PIN.json 2> $OUT/r2h_bench_pin$PIN.err; echo "pin $PIN exit $?"; grep hostprof $OUT/r2h_bench_pin$PIN.err
done
SDB_BENCH_HOSTPROF=1 SDB_BENCH_PIN_FILE=0 SDB_BENCH_PIPELINE=3 timeout 600 python bench.py --regime aged --no-configs --no-cpu-baseline --min-seconds 0.4 > $OUT/r2h_bench_pipe3.json 2> $OUT/r2h_bench_pipe3.err; echo "pipe3 exit $?"; grep hostprof $OUT/r2h_bench_pipe3.err
