#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_lifter.py tests/test_gpu_feeder.py -q > $O/r2q_pytest.log 2>&1; echo "tests exit $?"; tail -4 $O/r2q_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
for i in 1 2; do
$B > $O/r2q_train$i.json 2> $O/r2q_train$i.err; echo "train$i $?"; head -c 130 $O/r2q_train$i.json; echo
VP3D_NARROW=0 $B > $O/r2q_train_nonarrow$i.json 2> $O/r2q_train_nonarrow$i.err; echo "nonarrow$i $?"; head -c 130 $O/r2q_train_nonarrow$i.json; echo
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_act" -s 3 -c 3 -o $O/r2q_prof_bn -f python tools/ew_profile.py 27648 > $O/r2q_ncu_bn.log 2>&1; echo "ncu bn $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 200 --csv --log-file $O/r2q_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r2q_ncu_train.log 2>&1; echo "ncu launches $?"
