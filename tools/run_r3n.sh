#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_training.py tests/test_gpu_expand_fused.py tests/test_gpu_lifecycle.py -q -x > $O/r3n_pytest.log 2>&1; echo "tests exit $?"; tail -3 $O/r3n_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
for i in 1 2 3; do
$B > $O/r3n_train_$i.json 2>/dev/null; echo "train $i: $(python -c "import json;print(json.load(open('$O/r3n_train_$i.json'))['ms_per_step'])")"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bn_act_fwd" -s 24 -c 8 --csv --log-file $O/r3n_launches_bnfwd.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r3n_ncu.log 2>&1; echo "ncu $?"; grep bn_act_fwd $O/r3n_launches_bnfwd.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '
