#!/bin/bash
# lean epilogue: direct 32-byte global stores from registers against staged TMA stores
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2x_pytest.log 2>&1; echo "tests exit $?"; tail -6 $O/r2x_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
I="timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline --no-parity"
for i in 1 2 3; do
$I > $O/r2x_infer_direct_$i.json 2> $O/r2x_infer_direct_$i.err; echo "infer direct $i $?"; head -c 130 $O/r2x_infer_direct_$i.json; echo
VP3D_DIRECT_OUT=0 $I > $O/r2x_infer_tma_$i.json 2> $O/r2x_infer_tma_$i.err; echo "infer tma $i $?"; head -c 130 $O/r2x_infer_tma_$i.json; echo
done
for i in 1 2; do
$B > $O/r2x_train_direct_$i.json 2> $O/r2x_train_direct_$i.err; echo "train direct $i $?"; head -c 130 $O/r2x_train_direct_$i.json; echo
VP3D_DIRECT_OUT=0 $B > $O/r2x_train_tma_$i.json 2> $O/r2x_train_tma_$i.err; echo "train tma $i $?"; head -c 130 $O/r2x_train_tma_$i.json; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 130 --csv --log-file $O/r2x_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r2x_ncu_train.log 2>&1; echo "ncu launches $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_gemm|pack_rows" -s 33 -c 11 --csv --log-file $O/r2x_launches_infer.csv python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/r2x_ncu_infer.log 2>&1; echo "ncu infer launches $?"
