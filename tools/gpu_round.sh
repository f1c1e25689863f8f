#!/bin/bash
# One GPU-box visit: gpu tests, smoke, bench (N=1), launch list and one full ncu capture of the top kernels.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"; cat $OUT/${TAG}_bench.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_s2.json 2>/dev/null && \
SDB_BENCH_CUPROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
SDB_BENCH_CUPROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'dyn_step_kernel|struct3_kernel|ov_resolve_fused_kernel' -c 3 -o $OUT/${TAG}_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
python profiles/ncu_metrics.py $OUT/${TAG}_full_raw.csv > $OUT/${TAG}_full_summary.txt 2>&1; head -40 $OUT/${TAG}_full_summary.txt
