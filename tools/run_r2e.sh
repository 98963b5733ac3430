#!/bin/bash
# 2 GPUs: training headline with the new multi-GPU checks (params in sync, strong scaling, SyncBN)
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 30 --no-cpu-baseline --no-parity > $O/r2e_train_2gpu.json 2> $O/r2e_train_2gpu.err; echo "exit $?"; tail -c 2500 $O/r2e_train_2gpu.json; tail -5 $O/r2e_train_2gpu.err
