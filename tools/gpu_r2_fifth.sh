#!/bin/bash
# Round 2, fifth visit: whole GPU suite, bench (long pool, order path, segment floor), segment-floor sweep for configs[1].
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2e_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2e_pytest.log
tail -4 $OUT/r2e_pytest.log
timeout 1500 python bench.py > $OUT/r2e_bench.json 2> $OUT/r2e_bench.err; echo "bench exit $?"; tail -8 $OUT/r2e_bench.err
for SEG in 1 2 4 8; do
  SDB_SEG_MIN=$SEG timeout 600 python bench_configs.py --no-cpu-baseline > $OUT/r2e_cfg_seg$SEG.json 2> $OUT/r2e_cfg_seg$SEG.err; echo "cfg seg $SEG exit $?"
done
for T in 3 7 11; do
  SDB_COPY_THREADS=$T timeout 600 python bench.py --regime aged --no-configs --no-cpu-baseline --min-seconds 0.4 > $OUT/r2e_bench_copy$T.json 2> $OUT/r2e_bench_copy$T.err; echo "copy threads $T exit $?"
done
