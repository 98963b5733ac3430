#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r3o_pytest.log 2>&1; echo "tests exit $?"; tail -3 $O/r3o_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3o_smoke.log 2>&1; echo "smoke $?"; tail -2 $O/r3o_smoke.log
