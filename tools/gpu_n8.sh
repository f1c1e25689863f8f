#!/bin/bash
# Eight GPUs, configs[3], aged regime (the headline), frames that never run backwards.
OUT=gpurun_out; mkdir -p $OUT
SDB_SHARD_PROFILE=1 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --regime aged --no-configs --min-seconds 0.5 > $OUT/r2m_n8_bench.json 2> $OUT/r2m_n8_bench.err; echo "bench n8 exit $?"; grep "^\[rank [07]" $OUT/r2m_n8_bench.err; tail -1 $OUT/r2m_n8_bench.json | head -c 400
