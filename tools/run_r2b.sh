#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r2b_pytest_gpu.log 2>&1; echo "gpu tests exit $?"; tail -25 $O/r2b_pytest_gpu.log
timeout 600 python bench.py --mode train --steps 30 --no-cpu-baseline > $O/r2b_train.json 2> $O/r2b_train.err; echo "bench exit $?"; head -c 300 $O/r2b_train.json
echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file $O/r2b_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/r2b_ncu_train.log 2>&1; echo "ncu exit $?"
