#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2k_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2k_pytest.log
tail -4 $OUT/r2k_pytest.log
timeout 600 python bench_configs.py --no-cpu-baseline > $OUT/r2k_cfg.json 2> $OUT/r2k_cfg.err; echo "cfg exit $?"
timeout 600 python bench_configs.py --no-cpu-baseline > $OUT/r2k_cfg2.json 2> $OUT/r2k_cfg2.err; echo "cfg2 exit $?"
