python -m pytest tests/test_gpu_camera_loss.py -x -q 2>&1 | tail -2
for tune in 2,2304,3 2,1152,4 2,1536,4 2,768,4 3,1152,4 3,768,4 3,1536,3; do VP3D_PROJ_TUNE=$tune python tools/proj_probe.py 2>&1 | tail -3; done
