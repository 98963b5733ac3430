#!/bin/bash
O=gpurun_out
mkdir -p $O
for mode in dynamic static dynamic static; do
VP3D_DDP_SCHED=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r2g_train_2gpu_$mode.json 2> $O/r2g_train_2gpu_$mode.err; echo "$mode exit $?"; python - <<PY
import json
d=json.load(open('$O/r2g_train_2gpu_$mode.json'))
print('$mode', 'weak ms', d['ms_per_step'], 'strong ms', d['multi_gpu']['strong']['ms_per_step'], 'gemm', d['roofline']['gemm_ms_per_step'], d['multi_gpu']['params_in_sync'])
PY
done
