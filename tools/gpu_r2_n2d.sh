#!/bin/bash
# Two GPUs: the sliced e2e exchange with the library's async upload (was torch copy_), aged regime.
OUT=gpurun_out; mkdir -p $OUT
SDB_SHARD_PROFILE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --regime aged --no-configs --min-seconds 0.5 > $OUT/r2o_n2_bench.json 2> $OUT/r2o_n2_bench.err; echo "bench n2 exit $?"; grep "^\[rank" $OUT/r2o_n2_bench.err
SDB_BENCH_PIN_FILE=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --regime aged --no-configs --min-seconds 0.5 > $OUT/r2o_n2_bench_nopin.json 2> $OUT/r2o_n2_bench_nopin.err; echo "bench n2 nopin exit $?"
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/r2o_pytest_multi.log 2>&1; echo "multi exit $?"; tail -2 $OUT/r2o_pytest_multi.log
