#!/bin/bash
# PDL trigger placement in the long persistent kernels (A/B builds through VP3D_LIB_PATH): 1 GPU and 2 GPUs
O=gpurun_out
mkdir -p $O
L=$PWD/dynamic-camera-augmented-videopose3d_b200/lib
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
i=0
for lib in libvp3d_b200.so libvp3d_b200_late1.so libvp3d_b200_late2.so libvp3d_b200.so libvp3d_b200_late1.so libvp3d_b200_late2.so; do
i=$((i+1))
VP3D_LIB_PATH=$L/$lib timeout 300 python $B > $O/r3f_train1_$i.json 2>/dev/null; echo "1 GPU $lib: $(python -c "import json;print(json.load(open('$O/r3f_train1_$i.json'))['ms_per_step'])")"
done
i=0
for lib in libvp3d_b200.so libvp3d_b200_late1.so libvp3d_b200_late2.so libvp3d_b200_late1.so libvp3d_b200_late2.so; do
i=$((i+1))
VP3D_LIB_PATH=$L/$lib $T --master-port $((29580+i)) $B --gpus 2 > $O/r3f_train2_$i.json 2> $O/r3f_train2_$i.err; echo "2 GPUs $lib $?"; python - <<PY
import json
d=json.loads(open('$O/r3f_train2_$i.json').read().strip().splitlines()[-1])
print('  ms', d['ms_per_step'], 'strong', d['multi_gpu']['strong']['ms_per_step'], 'sync', d['multi_gpu']['params_in_sync'])
PY
done
