#!/bin/bash
# Round 2, third visit: compacted tracker slow path -- parity tests, bench in both regimes, ncu of the aged regime.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2c_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2c_pytest.log
tail -4 $OUT/r2c_pytest.log
for AGE in 1 10; do
  timeout 600 python bench.py --no-cpu-baseline --age-stride $AGE > $OUT/r2c_bench_age$AGE.json 2> $OUT/r2c_bench_age$AGE.err; echo "bench age $AGE exit $?"; cat $OUT/r2c_bench_age$AGE.json; tail -5 $OUT/r2c_bench_age$AGE.err
done
SDB_BENCH_CUPROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'dyn_step_kernel' -c 1 -o $OUT/r2c_aged_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --age-stride 10 > $OUT/r2c_ncu.log 2>&1; echo "ncu exit $?"
ncu -i $OUT/r2c_aged_full.ncu-rep --page raw --csv > $OUT/r2c_aged_full_raw.csv 2>/dev/null
python profiles/ncu_metrics.py $OUT/r2c_aged_full_raw.csv > $OUT/r2c_aged_full_summary.txt 2>&1; head -30 $OUT/r2c_aged_full_summary.txt
ncu -i $OUT/r2c_aged_full.ncu-rep --page source --csv > $OUT/r2c_aged_full_source.csv 2>/dev/null; wc -l $OUT/r2c_aged_full_source.csv
