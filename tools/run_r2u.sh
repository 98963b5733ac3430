#!/bin/bash
# lean epilogue with 16 epilogue warps for short contractions (EPI = 3): parity tests + same-box A/B
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2u_pytest.log 2>&1; echo "tests exit $?"; tail -6 $O/r2u_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
I="timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline --no-parity"
for i in 1 2; do
$B > $O/r2u_train_16_$i.json 2> $O/r2u_train_16_$i.err; echo "train lean16 $i $?"; head -c 130 $O/r2u_train_16_$i.json; echo
VP3D_LEAN16=0 $B > $O/r2u_train_8_$i.json 2> $O/r2u_train_8_$i.err; echo "train lean8 $i $?"; head -c 130 $O/r2u_train_8_$i.json; echo
$I > $O/r2u_infer_16_$i.json 2> $O/r2u_infer_16_$i.err; echo "infer lean16 $i $?"; head -c 130 $O/r2u_infer_16_$i.json; echo
VP3D_LEAN16=0 $I > $O/r2u_infer_8_$i.json 2> $O/r2u_infer_8_$i.err; echo "infer lean8 $i $?"; head -c 130 $O/r2u_infer_8_$i.json; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 130 --csv --log-file $O/r2u_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r2u_ncu_train.log 2>&1; echo "ncu launches $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_gemm|pack_rows" -s 33 -c 11 --csv --log-file $O/r2u_launches_infer.csv python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/r2u_ncu_infer.log 2>&1; echo "ncu infer launches $?"
