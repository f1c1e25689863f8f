#!/bin/bash
# A/B of environment knobs on the 200-origin bench: bash tools/gpu_ab.sh TAG "ENV1" "ENV2" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for E in "$@"; do
  i=$((i+1))
  env $E timeout 300 python bench.py --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$i.json"))
print("$E", "ms/step %.4f e2e %.4f dyn %.4f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["launch_ms"]), d["roofline"]["other_kernels_ms_per_step"])
PY
done
