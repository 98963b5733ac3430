#!/bin/bash
# ncu --set full with source of the expand-layer GEMM (K = 192: epilogue bound): inference (EPI = 0) and training
# (EPI = 1: BatchNorm + ReLU + dropout epilogue) plus the gated data gradient of block 1
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_pair_kernel" -c 1 -o $O/r2r_expand_infer -f python bench.py --mode infer --steps 1 --warmup 1 --seqs 16 --no-cpu-baseline --no-parity > $O/r2r_ncu_infer.log 2>&1; echo "ncu infer $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_pair_kernel<0, ., 1>" -c 2 -o $O/r2r_expand_train -f python bench.py --mode train --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-parity > $O/r2r_ncu_train.log 2>&1; echo "ncu train $?"
ls -la $O/r2r*
