#!/bin/bash
# Validation after the engine memory sizing change + struct wave knob at 25 origins.
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2n_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2n_pytest.log
tail -3 $OUT/r2n_pytest.log
A="--origins 25 --regime aged --no-configs --no-cpu-baseline --parity off --no-e2e --min-seconds 0.4"
for W in 0 6 12; do SDB_STRUCT_MIN_WAVES=$W timeout 300 python bench.py $A > $OUT/r2n_25_w$W.json 2> $OUT/r2n_25_w$W.err; echo "waves $W exit $?"; done
for W in 0 6 12; do python - <<PY
import json; d=json.load(open('$OUT/r2n_25_w$W.json')); r=d['regimes']['aged']; print('waves $W', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'dyn %.3f struct %.3f ov %.3f'%(r['roofline']['launch_ms'], r['roofline']['other_kernels_ms_per_step']['struct_kernel'], r['roofline']['other_kernels_ms_per_step']['overlap_bracket_resolve']))
PY
done
