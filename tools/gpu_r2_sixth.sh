#!/bin/bash
# Round 2, sixth visit: arena allocator, pinned GSD mapping, struct clear folded, batch finalize; side-stream threshold sweep.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2f_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2f_pytest.log
tail -4 $OUT/r2f_pytest.log
timeout 1500 python bench.py > $OUT/r2f_bench.json 2> $OUT/r2f_bench.err; echo "bench exit $?"; tail -8 $OUT/r2f_bench.err
SDB_SIDE_MIN_N=100000 timeout 600 python bench_configs.py --no-cpu-baseline > $OUT/r2f_cfg_noside.json 2> $OUT/r2f_cfg_noside.err; echo "cfg noside exit $?"
SDB_PIN_TRAJECTORY=0 SDB_BENCH_PIN_FILE=0 timeout 600 python bench.py --regime aged --no-configs --no-cpu-baseline --min-seconds 0.4 > $OUT/r2f_bench_nopin.json 2> $OUT/r2f_bench_nopin.err; echo "nopin exit $?"
