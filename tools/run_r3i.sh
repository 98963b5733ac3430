#!/bin/bash
# PDL trigger level 3 (HBM-bound passes keep their dependents back) against the default (2): 1 GPU and 2 GPUs
O=gpurun_out
mkdir -p $O
L=$PWD/dynamic-camera-augmented-videopose3d_b200/lib
B="bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
i=0
for lib in libvp3d_b200.so libvp3d_b200_late3.so libvp3d_b200.so libvp3d_b200_late3.so libvp3d_b200.so libvp3d_b200_late3.so; do
i=$((i+1))
VP3D_LIB_PATH=$L/$lib timeout 300 python $B > $O/r3i_train1_$i.json 2>/dev/null; echo "1 GPU $lib: $(python -c "import json;print(json.load(open('$O/r3i_train1_$i.json'))['ms_per_step'])")"
done
timeout 900 python -m pytest tests -m gpu -q -x > $O/r3i_pytest.log 2>&1; echo "tests exit $?"; tail -3 $O/r3i_pytest.log
