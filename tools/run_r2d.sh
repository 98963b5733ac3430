#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_expand_fused.py tests/test_gpu_training.py -q > $O/r2d_pytest.log 2>&1; echo "tests exit $?"; tail -8 $O/r2d_pytest.log
timeout 300 python tools/gram_probe.py > $O/r2d_gram.log 2>&1; cat $O/r2d_gram.log
for i in 1 2; do
VP3D_FIN_IN_GEMM=1 timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r2d_train_fin$i.json 2> $O/r2d_train_fin$i.err; echo "fin$i exit $?"; head -c 160 $O/r2d_train_fin$i.json; echo
VP3D_FIN_IN_GEMM=0 timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r2d_train_nofin$i.json 2> $O/r2d_train_nofin$i.err; echo "nofin$i exit $?"; head -c 160 $O/r2d_train_nofin$i.json; echo
done
