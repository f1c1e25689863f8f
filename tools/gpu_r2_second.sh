#!/bin/bash
# Round 2, second visit: the native engine under the whole GPU suite + bench in both regimes.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/r2b_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2b_pytest.log
tail -30 $OUT/r2b_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/r2b_smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/r2b_smoke.log
for AGE in 1 10; do
  timeout 600 python bench.py --no-cpu-baseline --age-stride $AGE > $OUT/r2b_bench_age$AGE.json 2> $OUT/r2b_bench_age$AGE.err; echo "bench age $AGE exit $?"; cat $OUT/r2b_bench_age$AGE.json; tail -5 $OUT/r2b_bench_age$AGE.err
done
timeout 600 python bench.py --no-cpu-baseline --origins 25 > $OUT/r2b_bench_25.json 2> $OUT/r2b_bench_25.err; echo "bench 25 origins exit $?"; cat $OUT/r2b_bench_25.json
SDB_BENCH_PYPROFILE=1 timeout 600 python bench.py --no-cpu-baseline --origins 25 > /dev/null 2> $OUT/r2b_pyprof_25.err; tail -60 $OUT/r2b_pyprof_25.err
