#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_temporal.py -q -x -k "stream" > $O/r2z_pytest.log 2>&1; echo "tests exit $?"; tail -15 $O/r2z_pytest.log
timeout 600 python tools/stream_probe.py > $O/r2z_stream_probe.json 2> $O/r2z_stream_probe.err; echo "probe $?"; cat $O/r2z_stream_probe.json; tail -3 $O/r2z_stream_probe.err
