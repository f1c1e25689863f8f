#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SDB_SHARD_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --regime aged --no-configs --min-seconds 0.5 > $OUT/r2q_n2_bench.json 2> $OUT/r2q_n2_bench.err; echo "bench n2 exit $?"; grep "^\[rank" $OUT/r2q_n2_bench.err
