#!/bin/bash
# Round 2: the bench with non-reversing frames (both regimes, parity, configs), then the evidence captures.
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python bench.py > $OUT/r2j_bench.json 2> $OUT/r2j_bench.err; echo "bench exit $?"; tail -5 $OUT/r2j_bench.err
timeout 400 python bench.py --origins 25 --no-configs --no-cpu-baseline --parity off --no-e2e --min-seconds 0.5 > $OUT/r2j_bench_25origins.json 2> $OUT/r2j_bench_25origins.err; echo "25 origins exit $?"
bash tools/gpu_r2_profiles.sh
