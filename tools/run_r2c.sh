#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r2c_pytest_gpu.log 2>&1; echo "gpu tests exit $?"; tail -25 $O/r2c_pytest_gpu.log
timeout 600 python bench.py --mode train --steps 30 --no-cpu-baseline > $O/r2c_train.json 2> $O/r2c_train.err; echo "bench exit $?"; head -c 300 $O/r2c_train.json; tail -5 $O/r2c_train.err
echo
VP3D_FIN_IN_GEMM=0 timeout 600 python bench.py --mode train --steps 30 --no-cpu-baseline --no-parity > $O/r2c_train_nofin.json 2> $O/r2c_train_nofin.err; echo "bench nofin exit $?"; head -c 300 $O/r2c_train_nofin.json
echo
timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline > $O/r2c_infer.json 2> $O/r2c_infer.err; echo "infer exit $?"; head -c 300 $O/r2c_infer.json; tail -3 $O/r2c_infer.err
echo
VP3D_RES_TMA=1 timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline > $O/r2c_infer_restma.json 2> $O/r2c_infer_restma.err; echo "infer restma exit $?"; head -c 300 $O/r2c_infer_restma.json; tail -3 $O/r2c_infer_restma.err
