#!/bin/bash
# 2-GPU training A/B: arena prefill and the lean epilogue under the peer exchange
O=gpurun_out
mkdir -p $O
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --mode train --steps 40 --no-cpu-baseline --no-parity"
i=0
for env in "X=1" "VP3D_PREFILL_ARENA=0" "VP3D_LEAN_EPI=0" "X=1" "VP3D_PREFILL_ARENA=0"; do
i=$((i+1))
env $env $T --master-port $((29520+i)) $B > $O/r3c_train2_$i.json 2> $O/r3c_train2_$i.err; echo "$env $?"; python - <<PY
import json
d=json.loads(open('$O/r3c_train2_$i.json').read().strip().splitlines()[-1])
print('  ms', d['ms_per_step'], 'strong', d['multi_gpu']['strong']['ms_per_step'])
PY
done
timeout 300 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r3c_train1.json 2>/dev/null; head -c 120 $O/r3c_train1.json; echo
