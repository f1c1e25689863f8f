#!/bin/bash
# Full ncu capture of kernels matching REGEX inside the timed region of a short bench run.
# Usage: bash tools/gpu_ncu.sh TAG REGEX COUNT [bench args...]
TAG=$1; REGEX=$2; COUNT=$3; shift 3
OUT=gpurun_out
mkdir -p $OUT
SDB_BENCH_CUPROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"$REGEX" -c $COUNT -o $OUT/${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_ncu.log 2>&1; echo "ncu exit $?"
ncu -i $OUT/${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
python profiles/ncu_metrics.py $OUT/${TAG}_raw.csv > $OUT/${TAG}_summary.txt 2>&1; cat $OUT/${TAG}_summary.txt
