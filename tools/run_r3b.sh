#!/bin/bash
# the driver's multi-GPU command (mode all) on 2 GPUs + the reference arm under torchrun
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r3b_bench_2gpu.json 2> $O/r3b_bench_2gpu.err; echo "bench 2gpu $?"; tail -c 400 $O/r3b_bench_2gpu.json; tail -5 $O/r3b_bench_2gpu.err
timeout 600 python -m pytest tests/test_gpu_peer_allreduce.py -q > $O/r3b_pytest_peer.log 2>&1; echo "peer tests $?"; tail -3 $O/r3b_pytest_peer.log
