#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2v_pytest.log 2>&1; echo "tests exit $?"; tail -6 $O/r2v_pytest.log
