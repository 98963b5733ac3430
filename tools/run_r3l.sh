#!/bin/bash
# lean forward BatchNorm pass; stored keep bits on / off
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r3l_pytest.log 2>&1; echo "tests exit $?"; tail -5 $O/r3l_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
for i in 1 2 3; do
$B > $O/r3l_train_nomask_$i.json 2>/dev/null; echo "no mask $i: $(python -c "import json;print(json.load(open('$O/r3l_train_nomask_$i.json'))['ms_per_step'])")"
VP3D_BN_MASK=1 $B > $O/r3l_train_mask_$i.json 2>/dev/null; echo "mask $i: $(python -c "import json;print(json.load(open('$O/r3l_train_mask_$i.json'))['ms_per_step'])")"
done
VP3D_BN_MASK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 130 --csv --log-file $O/r3l_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r3l_ncu_train.log 2>&1; echo "ncu launches $?"
