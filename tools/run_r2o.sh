#!/bin/bash
# 8-GPU check of the peer-memory gradient exchange and the training scaling numbers.
O=gpurun_out
mkdir -p $O
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="bench.py --gpus $N --mode train --steps 40 --no-cpu-baseline --no-parity"
timeout 300 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r2o_train_1gpu.json 2> $O/r2o_train_1gpu.err; echo "1gpu $?"; head -c 130 $O/r2o_train_1gpu.json; echo
timeout 300 $TR --master-port 29611 tools/ddp_peer_check.py > $O/r2o_peer_check_$N.json 2> $O/r2o_peer_check_$N.err; echo "check exit $?"; cat $O/r2o_peer_check_$N.json; tail -3 $O/r2o_peer_check_$N.err
timeout 400 $TR --master-port 29612 $B --exchange peer > $O/r2o_train_peer8_$N.json 2> $O/r2o_train_peer8_$N.err; echo "peer ctas8 $?"; head -c 130 $O/r2o_train_peer8_$N.json; echo; tail -2 $O/r2o_train_peer8_$N.err
VP3D_DDP_CTAS=4 timeout 400 $TR --master-port 29613 $B --exchange peer > $O/r2o_train_peer4_$N.json 2> $O/r2o_train_peer4_$N.err; echo "peer ctas4 $?"; head -c 130 $O/r2o_train_peer4_$N.json; echo
timeout 400 $TR --master-port 29614 $B --exchange nccl > $O/r2o_train_nccl_$N.json 2> $O/r2o_train_nccl_$N.err; echo "nccl $?"; head -c 130 $O/r2o_train_nccl_$N.json; echo
timeout 400 $TR --master-port 29615 bench.py --gpus $N --mode c4 --steps 10 --no-cpu-baseline --no-parity > $O/r2o_c4_$N.json 2> $O/r2o_c4_$N.err; echo "c4 $?"; head -c 200 $O/r2o_c4_$N.json; echo; tail -2 $O/r2o_c4_$N.err
