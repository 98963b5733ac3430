#!/bin/bash
# 2-GPU training A/B: optimiser update inside the backward / PDL under the peer exchange
O=gpurun_out
mkdir -p $O
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --mode train --steps 40 --no-cpu-baseline --no-parity"
i=0
for cfg in "X=1|" "X=1|--no-update-in-backward" "VP3D_PDL=0|" "VP3D_PDL=0|--no-update-in-backward" "VP3D_DDP_CTAS=4|" "VP3D_DDP_CTAS=16|" "X=1|--no-update-in-backward"; do
i=$((i+1))
env=${cfg%%|*}; flag=${cfg##*|}
env $env $T --master-port $((29540+i)) $B $flag > $O/r3d_train2_$i.json 2> $O/r3d_train2_$i.err; echo "$env $flag $?"; python - <<PY
import json
d=json.loads(open('$O/r3d_train2_$i.json').read().strip().splitlines()[-1])
print('  ms', d['ms_per_step'], 'strong', d['multi_gpu']['strong']['ms_per_step'])
PY
done
