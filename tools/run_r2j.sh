#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r2j_$tag.json 2> $O/r2j_$tag.err; echo "$tag exit $?"; python - <<PY
import json
try:
    d=json.load(open('$O/r2j_$tag.json')); r=d['roofline']
    print('$tag', 'weak ms', round(d['ms_per_step'],4), 'strong ms', round(d['multi_gpu']['strong']['ms_per_step'],4), 'conv', round(r['conv_fwd_dgrad_ms'],3), 'wgrad', round(r['wgrad_ms'],3), d['config']['grad_exchange'][:90])
except Exception as e:
    print('$tag failed', e)
PY
}
run static VP3D_DDP_SCHED=static
run onebucket VP3D_DDP_SCHED=static VP3D_DDP_LARGE_BYTES=1099511627776
run static_nvls VP3D_DDP_SCHED=static NCCL_ALGO=NVLS
run static_ll128 VP3D_DDP_SCHED=static NCCL_PROTO=LL128
run static_simple VP3D_DDP_SCHED=static NCCL_PROTO=Simple
