#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_lifecycle.py tests/test_gpu_expand_fused.py -q -x > $O/r2i_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2i_tests.log
run() { tag=$1; shift; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r2i_$tag.json 2> $O/r2i_$tag.err; echo "$tag exit $?"; python - <<PY
import json
try:
    d=json.load(open('$O/r2i_$tag.json')); r=d['roofline']
    print('$tag', 'weak ms', round(d['ms_per_step'],4), 'strong ms', round(d['multi_gpu']['strong']['ms_per_step'],4), 'conv', round(r['conv_fwd_dgrad_ms'],3), 'wgrad', round(r['wgrad_ms'],3), d['multi_gpu']['params_in_sync'])
except Exception as e:
    print('$tag failed', e)
PY
}
run dyn VP3D_DDP_SCHED=dynamic
run static VP3D_DDP_SCHED=static
run dyn_cta8 VP3D_DDP_SCHED=dynamic NCCL_MAX_CTAS=8
run dyn_cta16 VP3D_DDP_SCHED=dynamic NCCL_MAX_CTAS=16
