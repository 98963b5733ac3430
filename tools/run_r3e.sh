#!/bin/bash
# 2-GPU training: all-reduce kernel without early release of its dependents
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_allreduce.py -q > $O/r3e_pytest_peer.log 2>&1; echo "peer tests $?"; tail -3 $O/r3e_pytest_peer.log
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --mode train --steps 40 --no-cpu-baseline --no-parity"
i=0
for cfg in "X=1|" "VP3D_PDL=0|" "X=1|--no-update-in-backward" "X=1|" "VP3D_PDL=0|" "VP3D_DDP_CTAS=16|"; do
i=$((i+1))
env=${cfg%%|*}; flag=${cfg##*|}
env $env $T --master-port $((29560+i)) $B $flag > $O/r3e_train2_$i.json 2> $O/r3e_train2_$i.err; echo "$env $flag $?"; python - <<PY
import json
d=json.loads(open('$O/r3e_train2_$i.json').read().strip().splitlines()[-1])
print('  ms', d['ms_per_step'], 'strong', d['multi_gpu']['strong']['ms_per_step'], 'sync', d['multi_gpu']['params_in_sync'])
PY
done
timeout 300 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r3e_train1.json 2>/dev/null; head -c 120 $O/r3e_train1.json; echo
