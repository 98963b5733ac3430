#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_expand_fused.py -q -x -k "dynamic or side or gated" > $O/r2f_dyn.log 2>&1; echo "dyn tests exit $?"; tail -12 $O/r2f_dyn.log
timeout 900 python -m pytest tests -m gpu -q > $O/r2f_pytest.log 2>&1; echo "all tests exit $?"; tail -6 $O/r2f_pytest.log
timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline > $O/r2f_infer_static.json 2> $O/r2f_infer_static.err; echo "infer static $?"; head -c 250 $O/r2f_infer_static.json; echo
VP3D_SCHED=dynamic timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline > $O/r2f_infer_dyn.json 2> $O/r2f_infer_dyn.err; echo "infer dyn $?"; head -c 250 $O/r2f_infer_dyn.json; echo; tail -3 $O/r2f_infer_dyn.err
VP3D_SCHED=dynamic timeout 600 python bench.py --mode train --steps 30 --no-cpu-baseline --no-parity > $O/r2f_train_dyn.json 2> $O/r2f_train_dyn.err; echo "train dyn $?"; head -c 200 $O/r2f_train_dyn.json; echo
timeout 600 python bench.py --mode train --steps 30 --no-cpu-baseline --no-parity > $O/r2f_train_static.json 2> $O/r2f_train_static.err; echo "train static $?"; head -c 200 $O/r2f_train_static.json; echo
