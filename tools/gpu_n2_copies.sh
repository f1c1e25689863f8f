#!/bin/bash
# Two GPUs: tiny copies (descriptor upload, peer row write) on the SMs vs on the copy engines, end-to-end.
OUT=gpurun_out; mkdir -p $OUT
SDB_SMALL_COPIES=kernel timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_tests.py tests/test_gpu_cabi.py tests/test_gpu_multi.py tests/test_gpu_process_file.py -x -q > $OUT/r2r_pytest.log 2>&1; echo "pytest (SM copies) exit $?"; tail -2 $OUT/r2r_pytest.log
A="--gpus 2 --regime aged --no-configs --min-seconds 0.4"
SDB_SMALL_COPIES=kernel timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py $A > $OUT/r2r_n2_smcopies.json 2> $OUT/r2r_n2_smcopies.err; echo "sm copies exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py $A > $OUT/r2r_n2_default.json 2> $OUT/r2r_n2_default.err; echo "default exit $?"
