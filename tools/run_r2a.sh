#!/bin/bash
# first GPU contact of round 2: new fused-expand tests, training / conv tests, quick training bench (fused on / off)
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_expand_fused.py -x -q -s > $O/r2a_expand.log 2>&1; echo "expand tests exit $?"; tail -15 $O/r2a_expand.log
timeout 1200 python -m pytest tests/test_gpu_training.py tests/test_gpu_conv_gemm.py -x -q > $O/r2a_train.log 2>&1; echo "training tests exit $?"; tail -5 $O/r2a_train.log
timeout 600 python bench.py --mode train --steps 30 --no-cpu-baseline > $O/r2a_train_fused.json 2> $O/r2a_train_fused.err; echo "bench fused exit $?"; head -c 1500 $O/r2a_train_fused.json
VP3D_FUSED_EXPAND=0 timeout 600 python bench.py --mode train --steps 30 --no-cpu-baseline > $O/r2a_train_unfused.json 2> $O/r2a_train_unfused.err; echo "bench unfused exit $?"; head -c 600 $O/r2a_train_unfused.json
