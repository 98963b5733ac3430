#!/bin/bash
# gate of the data gradient as compare-to-mask on the packed pairs; ncu source profiles of the lean expand launches
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_expand_fused.py tests/test_gpu_training.py tests/test_gpu_lifter.py tests/test_gpu_conv_gemm.py -q -x > $O/r2t_pytest.log 2>&1; echo "tests exit $?"; tail -6 $O/r2t_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
for i in 1 2; do
$B > $O/r2t_train$i.json 2> $O/r2t_train$i.err; echo "train$i $?"; head -c 130 $O/r2t_train$i.json; echo
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_pair_kernel" -c 1 -o $O/r2t_expand_infer -f python bench.py --mode infer --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/r2t_ncu_infer.log 2>&1; echo "ncu infer $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_pair_kernel" -c 1 -o $O/r2t_expand_train -f python bench.py --mode train --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-parity > $O/r2t_ncu_train.log 2>&1; echo "ncu train $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 130 --csv --log-file $O/r2t_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r2t_ncu_train2.log 2>&1; echo "ncu launches $?"
