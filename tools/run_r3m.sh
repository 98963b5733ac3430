#!/bin/bash
# final build: GPU tests, smoke, the two headline bench lines, launch list of a training step
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r3m_pytest.log 2>&1; echo "tests exit $?"; tail -3 $O/r3m_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3m_smoke.log 2>&1; echo "smoke $?"; tail -2 $O/r3m_smoke.log
timeout 900 python bench.py > $O/r3m_bench_1gpu.json 2> $O/r3m_bench_1gpu.err; echo "bench $?"; head -c 200 $O/r3m_bench_1gpu.json; echo
timeout 600 python bench.py --mode train > $O/r3m_bench_train_1gpu.json 2> $O/r3m_bench_train_1gpu.err; echo "bench train $?"; head -c 160 $O/r3m_bench_train_1gpu.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r3m_bench_reference_arm.json 2>/dev/null; echo "ref arm $?"; head -c 200 $O/r3m_bench_reference_arm.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 220 --csv --log-file $O/r3m_launches_train.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-parity > $O/r3m_ncu_train.log 2>&1; echo "ncu launches $?"
