#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r2l_pytest.log 2>&1; echo "tests exit $?"; tail -5 $O/r2l_pytest.log
timeout 300 python tools/gram_probe.py > $O/r2l_gram.log 2>&1; cat $O/r2l_gram.log
for i in 1 2; do
timeout 600 python bench.py --mode infer --steps 20 --no-cpu-baseline --no-parity > $O/r2l_infer$i.json 2> $O/r2l_infer$i.err; echo "infer$i $?"; head -c 230 $O/r2l_infer$i.json; echo
timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r2l_train$i.json 2> $O/r2l_train$i.err; echo "train$i $?"; head -c 200 $O/r2l_train$i.json; echo
done
