#!/bin/bash
# backward's zero arena filled during the forward (side stream): tests + same-box A/B
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r3a_pytest.log 2>&1; echo "tests exit $?"; tail -5 $O/r3a_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
for i in 1 2 3; do
$B > $O/r3a_train_prefill_$i.json 2> $O/r3a_train_prefill_$i.err; echo "prefill $i $?"; head -c 130 $O/r3a_train_prefill_$i.json; echo
VP3D_PREFILL_ARENA=0 $B > $O/r3a_train_base_$i.json 2> $O/r3a_train_base_$i.err; echo "base $i $?"; head -c 130 $O/r3a_train_base_$i.json; echo
done
