#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r3g_pytest.log 2>&1; echo "tests exit $?"; tail -4 $O/r3g_pytest.log
for i in 1 2; do
timeout 300 python bench.py --mode infer --steps 30 --no-cpu-baseline --no-parity > $O/r3g_infer_$i.json 2>/dev/null; echo "infer $i: $(python -c "import json;print(json.load(open('$O/r3g_infer_$i.json'))['ms_per_step'])")"
timeout 300 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity > $O/r3g_train_$i.json 2>/dev/null; echo "train $i: $(python -c "import json;print(json.load(open('$O/r3g_train_$i.json'))['ms_per_step'])")"
done
timeout 300 python tools/stream_probe.py > $O/r3g_stream.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/r3g_stream.json'))
print({k:round(v['ms_device'],4) for k,v in d.items() if 'ms_device' in v})"
