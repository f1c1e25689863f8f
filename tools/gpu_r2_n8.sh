#!/bin/bash
# Round 2, eight GPUs: scaling of configs[3] and configs[4] at its real shape.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/r2n8_gpus.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
SDB_SHARD_PROFILE=1 timeout 900 $RUN --master-port 29521 bench.py --gpus 8 > $OUT/r2n8_bench.json 2> $OUT/r2n8_bench.err; echo "bench n8 exit $?"; grep "rank 0\|rank 3" $OUT/r2n8_bench.err | grep -v Warn | tail -8; tail -1 $OUT/r2n8_bench.json | head -c 1500; echo
timeout 900 $RUN --master-port 29522 bench.py --gpus 8 --molecules 4000000 --origins 500 --regime aged --no-configs > $OUT/r2n8_config4.json 2> $OUT/r2n8_config4.err; echo "config4 n8 exit $?"; tail -3 $OUT/r2n8_config4.err; tail -1 $OUT/r2n8_config4.json | head -c 1500; echo
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -s > $OUT/r2n8_pytest_multi.log 2>&1; echo "multi pytest exit $?"; tail -3 $OUT/r2n8_pytest_multi.log
