#!/bin/bash
# Quick visit: parity tests of the hot path, bench at 200 and 25 origins (the per-rank load of 8 GPUs), launch lists.
TAG=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
SDB_SHARD_PROFILE=1 timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench200.json 2> $OUT/${TAG}_bench200.err; echo "bench exit $?"; cat $OUT/${TAG}_bench200.json; grep "host enqueue" $OUT/${TAG}_bench200.err
SDB_SHARD_PROFILE=1 timeout 600 python bench.py --no-cpu-baseline --origins 25 --steps 40 --warmup 60 > $OUT/${TAG}_bench25.json 2> $OUT/${TAG}_bench25.err; echo "bench25 exit $?"; cat $OUT/${TAG}_bench25.json; grep "host enqueue" $OUT/${TAG}_bench25.err
for O in 200 25; do
SDB_BENCH_CUPROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches$O.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --origins $O > $OUT/${TAG}_ncu_launches$O.log 2>&1; echo "ncu launches $O exit $?"
done
