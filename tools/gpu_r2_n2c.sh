#!/bin/bash
# Final two-GPU record: the default bench (both regimes, GSD e2e, parity of two origins vs the oracle).
OUT=gpurun_out; mkdir -p $OUT
SDB_SHARD_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --parity on > $OUT/r2l_n2_bench.json 2> $OUT/r2l_n2_bench.err; echo "bench n2 exit $?"; grep "^\[rank" $OUT/r2l_n2_bench.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --molecules 4000000 --origins 125 --regime aged --no-configs --min-seconds 0.5 > $OUT/r2l_n2_config4_share.json 2> $OUT/r2l_n2_config4_share.err; echo "config4 share exit $?"; tail -2 $OUT/r2l_n2_config4_share.err | cut -c1-300
