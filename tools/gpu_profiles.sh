#!/bin/bash
# Round 2 evidence visit (1 GPU): launch list of the timed regions, full ncu captures of the step kernels in both
# regimes and of the neighbour search, SASS opcode histogram.  Numbers printed under ncu are never bench values.
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e --min-seconds 0"
for REG in aged early; do
  SDB_BENCH_CUPROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/r2p_launches_$REG.csv $B --regime $REG > $OUT/r2p_launches_$REG.log 2>&1; echo "launch list $REG exit $?"
  SDB_BENCH_CUPROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:'dyn_step_kernel|struct3_kernel|ov_resolve_fused_kernel' -c 3 -o $OUT/r2p_full_$REG -f $B --regime $REG > $OUT/r2p_full_$REG.log 2>&1; echo "ncu full $REG exit $?"
  ncu -i $OUT/r2p_full_$REG.ncu-rep --page raw --csv > $OUT/r2p_full_${REG}_raw.csv 2>/dev/null
  python profiles/ncu_metrics.py $OUT/r2p_full_${REG}_raw.csv > $OUT/r2p_full_${REG}_summary.txt 2>&1
  ncu -i $OUT/r2p_full_$REG.ncu-rep --page source --csv --kernel-name regex:dyn_step > $OUT/r2p_full_${REG}_source.csv 2>/dev/null
  python profiles/sass_hot.py $OUT/r2p_full_${REG}_source.csv > $OUT/r2p_sass_${REG}.txt 2>&1
done
head -20 $OUT/r2p_full_aged_summary.txt
# neighbour search (configs[2]): 32 k and 1 M molecules
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'nb_search2_kernel' -c 2 -o $OUT/r2p_nb_search -f \
    python tools/order_bench.py > $OUT/r2p_nb_search.log 2>&1; echo "ncu nb_search exit $?"
ncu -i $OUT/r2p_nb_search.ncu-rep --page raw --csv > $OUT/r2p_nb_search_raw.csv 2>/dev/null
python profiles/ncu_metrics.py $OUT/r2p_nb_search_raw.csv > $OUT/r2p_nb_search_summary.txt 2>&1; head -40 $OUT/r2p_nb_search_summary.txt
rm -f $OUT/*.ncu-rep
