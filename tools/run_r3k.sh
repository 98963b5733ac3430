#!/bin/bash
# BatchNorm backward with stored keep bits: tests, same-box A/B, launch list
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r3k_pytest.log 2>&1; echo "tests exit $?"; tail -5 $O/r3k_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
for i in 1 2 3; do
$B > $O/r3k_train_mask_$i.json 2>/dev/null; echo "mask $i: $(python -c "import json;print(json.load(open('$O/r3k_train_mask_$i.json'))['ms_per_step'])")"
VP3D_BN_MASK=0 $B > $O/r3k_train_base_$i.json 2>/dev/null; echo "base $i: $(python -c "import json;print(json.load(open('$O/r3k_train_base_$i.json'))['ms_per_step'])")"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 130 --csv --log-file $O/r3k_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r3k_ncu_train.log 2>&1; echo "ncu launches $?"
