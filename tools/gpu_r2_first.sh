#!/bin/bash
# Round 2, first visit: the new sized parity tests + the aged regime of the round-1 kernels.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/r2a_gpu.txt
nproc > $OUT/r2a_host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/r2a_host.txt; free -g >> $OUT/r2a_host.txt; df -h /dev/shm /tmp >> $OUT/r2a_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/r2a_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2a_pytest.log
tail -25 $OUT/r2a_pytest.log
for AGE in 1 10; do
  timeout 600 python bench.py --no-cpu-baseline --age-stride $AGE > $OUT/r2a_bench_age$AGE.json 2> $OUT/r2a_bench_age$AGE.err; echo "bench age $AGE exit $?"; cat $OUT/r2a_bench_age$AGE.json
done
SDB_BENCH_CUPROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'dyn_step_kernel|struct3_kernel|ov_resolve_fused_kernel' -c 3 -o $OUT/r2a_aged_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --age-stride 10 > $OUT/r2a_ncu.log 2>&1; echo "ncu exit $?"
ncu -i $OUT/r2a_aged_full.ncu-rep --page raw --csv > $OUT/r2a_aged_full_raw.csv 2>/dev/null
python profiles/ncu_metrics.py $OUT/r2a_aged_full_raw.csv > $OUT/r2a_aged_full_summary.txt 2>&1; head -60 $OUT/r2a_aged_full_summary.txt
