#!/bin/bash
# Final single-GPU visit of round 2: the whole GPU suite, smoke, the default bench line.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2z_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2z_pytest.log
tail -3 $OUT/r2z_pytest.log
timeout 200 python __graft_entry__.py --smoke > $OUT/r2z_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/r2z_smoke.log
timeout 900 python bench.py > $OUT/r2z_bench.json 2> $OUT/r2z_bench.err; echo "bench exit $?"
