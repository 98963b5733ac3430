#!/bin/bash
O=gpurun_out
mkdir -p $O
I="timeout 600 python bench.py --mode infer --steps 30 --no-cpu-baseline --no-parity"
for i in 1 2 3 4; do
$I > $O/r2y_infer_direct_$i.json 2> $O/r2y_infer_direct_$i.err; echo "infer direct $i $?"; head -c 130 $O/r2y_infer_direct_$i.json; echo
VP3D_DIRECT_OUT=0 $I > $O/r2y_infer_tma_$i.json 2> $O/r2y_infer_tma_$i.err; echo "infer tma $i $?"; head -c 130 $O/r2y_infer_tma_$i.json; echo
done
