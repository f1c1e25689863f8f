#!/bin/bash
# Where does the multi-GPU step lose time at 25 origins per rank?  (2 GPUs x 25 origins = the per-rank load of 8 GPUs)
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
A="--gpus 2 --origins 50 --regime early --no-e2e --no-configs --parity off --min-seconds 0.3"
SDB_SHARD_PROFILE=1 timeout 300 $RUN --master-port 29531 bench.py $A > $OUT/r2i_n2_direct.json 2> $OUT/r2i_n2_direct.err; echo "direct exit $?"; grep "^\[rank" $OUT/r2i_n2_direct.err
SDB_SHARD_PROFILE=1 SDB_BENCH_DIRECT_ROWS=0 timeout 300 $RUN --master-port 29532 bench.py $A > $OUT/r2i_n2_torchrows.json 2> $OUT/r2i_n2_torchrows.err; echo "torch rows exit $?"; grep "^\[rank" $OUT/r2i_n2_torchrows.err
timeout 300 $RUN --master-port 29533 bench.py $A > $OUT/r2i_n2_noprofile.json 2> $OUT/r2i_n2_noprofile.err; echo "no profile exit $?"
SDB_BENCH_HOSTPROF=1 timeout 300 python bench.py --origins 25 --regime early --no-e2e --no-configs --parity off --no-cpu-baseline --min-seconds 0.3 > $OUT/r2i_n1_25.json 2> $OUT/r2i_n1_25.err; echo "n1 exit $?"; grep hostprof $OUT/r2i_n1_25.err
for f in r2i_n2_direct r2i_n2_torchrows r2i_n2_noprofile r2i_n1_25; do tail -1 $OUT/$f.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['regimes']['early']; print('$f', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'sust %.4f'%r['sustained']['ms_per_step'], 'dyn %.3f struct %.3f ov %.3f'%(r['roofline']['launch_ms'], r['roofline']['other_kernels_ms_per_step']['struct_kernel'], r['roofline']['other_kernels_ms_per_step']['overlap_bracket_resolve']), 'host %.3f'%r['host_enqueue_ms_per_step'], r['regions']['min_ms'], r['regions']['max_ms'])"; done
