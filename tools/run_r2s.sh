#!/bin/bash
# lean epilogue (EPI = 2) of the pair kernel: parity tests, then same-box A/B of the training and the inference step
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2s_pytest.log 2>&1; echo "tests exit $?"; tail -6 $O/r2s_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
I="timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline --no-parity"
for i in 1 2; do
$B > $O/r2s_train_lean$i.json 2> $O/r2s_train_lean$i.err; echo "train lean$i $?"; head -c 130 $O/r2s_train_lean$i.json; echo
VP3D_LEAN_EPI=0 $B > $O/r2s_train_base$i.json 2> $O/r2s_train_base$i.err; echo "train base$i $?"; head -c 130 $O/r2s_train_base$i.json; echo
$I > $O/r2s_infer_lean$i.json 2> $O/r2s_infer_lean$i.err; echo "infer lean$i $?"; head -c 130 $O/r2s_infer_lean$i.json; echo
VP3D_LEAN_EPI=0 $I > $O/r2s_infer_base$i.json 2> $O/r2s_infer_base$i.err; echo "infer base$i $?"; head -c 130 $O/r2s_infer_base$i.json; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 130 --csv --log-file $O/r2s_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r2s_ncu_train.log 2>&1; echo "ncu launches $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_gemm|pack_rows" -s 33 -c 11 --csv --log-file $O/r2s_launches_infer.csv python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/r2s_ncu_infer.log 2>&1; echo "ncu infer launches $?"
