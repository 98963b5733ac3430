#!/bin/bash
# 2-GPU check of the peer-memory gradient exchange: unit test (one device), multicast / peer paths against NCCL under
# torchrun, then the training bench with both exchanges.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_peer_allreduce.py -q -x > $O/r2m_pytest.log 2>&1; echo "unit exit $?"; tail -3 $O/r2m_pytest.log
timeout 300 $TR --master-port 29511 tools/ddp_peer_check.py > $O/r2m_peer_check.json 2> $O/r2m_peer_check.err; echo "check exit $?"; cat $O/r2m_peer_check.json; tail -5 $O/r2m_peer_check.err
port=29520
for ex in peer nccl peer nccl; do
  port=$((port+1))
  timeout 600 $TR --master-port $port bench.py --gpus 2 --mode train --steps 40 --no-cpu-baseline --no-parity --exchange $ex > $O/r2m_train_${ex}_$port.json 2> $O/r2m_train_${ex}_$port.err
  echo "train $ex exit $?"; head -c 300 $O/r2m_train_${ex}_$port.json; echo; tail -3 $O/r2m_train_${ex}_$port.err
done
nvidia-smi topo -m > $O/r2m_topo.txt 2>&1
