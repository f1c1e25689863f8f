#!/bin/bash
# Round 2, two GPUs: sharded parity test + bench at N=2 (with the oracle parity check of two aged origins).
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/r2n2_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -s > $OUT/r2n2_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2n2_pytest.log
tail -6 $OUT/r2n2_pytest.log
SDB_SHARD_PROFILE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --parity on > $OUT/r2n2_bench.json 2> $OUT/r2n2_bench.err; echo "bench n2 exit $?"; tail -12 $OUT/r2n2_bench.err
timeout 900 python bench.py --origins 100 --no-configs --no-cpu-baseline > $OUT/r2n2_bench_1gpu_100origins.json 2> $OUT/r2n2_bench_1gpu_100origins.err; echo "bench 1gpu 100 origins exit $?"
