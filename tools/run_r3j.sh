#!/bin/bash
# PDL tail trigger in the persistent GEMMs against the default (no trigger): 1 GPU train + infer, 2 GPUs train
O=gpurun_out
mkdir -p $O
L=$PWD/dynamic-camera-augmented-videopose3d_b200/lib
B="bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
I="bench.py --mode infer --steps 20 --no-cpu-baseline --no-parity"
i=0
for lib in libvp3d_b200.so libvp3d_b200_tail.so libvp3d_b200.so libvp3d_b200_tail.so libvp3d_b200.so libvp3d_b200_tail.so; do
i=$((i+1))
VP3D_LIB_PATH=$L/$lib timeout 300 python $B > $O/r3j_train1_$i.json 2>/dev/null; echo "1 GPU train $lib: $(python -c "import json;print(json.load(open('$O/r3j_train1_$i.json'))['ms_per_step'])")"
done
for lib in libvp3d_b200.so libvp3d_b200_tail.so libvp3d_b200.so libvp3d_b200_tail.so; do
i=$((i+1))
VP3D_LIB_PATH=$L/$lib timeout 300 python $I > $O/r3j_infer1_$i.json 2>/dev/null; echo "1 GPU infer $lib: $(python -c "import json;print(json.load(open('$O/r3j_infer1_$i.json'))['ms_per_step'])")"
done
VP3D_LIB_PATH=$L/libvp3d_b200_tail.so timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_conv_gemm.py tests/test_gpu_temporal.py -m gpu -q -x > $O/r3j_pytest.log 2>&1; echo "tests (tail lib) exit $?"; tail -3 $O/r3j_pytest.log
