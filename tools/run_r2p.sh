#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r2p_pytest.log 2>&1; echo "tests exit $?"; tail -6 $O/r2p_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
$B > $O/r2p_train.json 2> $O/r2p_train.err; echo "train $?"; head -c 130 $O/r2p_train.json; echo
