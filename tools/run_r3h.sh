#!/bin/bash
# 8 GPUs: training line with the peer exchange (final build), the driver's all-mode command, one GPU of the same box
O=gpurun_out
mkdir -p $O
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29601 bench.py --gpus 8 --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r3h_train8.json 2> $O/r3h_train8.err; echo "train8 $?"
python - <<PY
import json
d=json.loads(open('$O/r3h_train8.json').read().strip().splitlines()[-1])
print('  ms', d['ms_per_step'], 'value', d['value'], 'strong', d['multi_gpu']['strong']['ms_per_step'], 'sync', d['multi_gpu']['params_in_sync'], 'syncbn', d['multi_gpu']['syncbn']['loss_rel_delta'])
PY
timeout 200 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r3h_train1.json 2>/dev/null; echo "train1: $(python -c "import json;print(json.load(open('$O/r3h_train1.json'))['ms_per_step'])")"
$T --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r3h_all8.json 2> $O/r3h_all8.err; echo "all8 $?"; tail -c 300 $O/r3h_all8.json
