"""Pivot an `ncu --metrics ... --csv --log-file X` launch list into one row per launch, and optionally merge the average
DRAM bytes per launch of the kernels matching a regex into profiles/r01_traffic.json.
Usage: python tools/ncu_metrics_table.py in.csv out.csv [traffic.json key regex]"""
import csv
import json
import os
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, mi, ui, vi = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Unit'), H.index('Metric Value')
launches, order, metrics = {}, [], []
for r in data:
    if r[0] not in launches:
        launches[r[0]] = {'kernel': r[ki].split('(')[0].replace('void ', '').replace('vp3d::', '')}
        order.append(r[0])
    name = '%s [%s]' % (r[mi], r[ui])
    if name not in metrics:
        metrics.append(name)
    launches[r[0]][name] = float(r[vi].replace(',', ''))
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['id', 'kernel'] + metrics)
    for i in order:
        w.writerow([i, launches[i]['kernel']] + [launches[i].get(m, '') for m in metrics])
if len(sys.argv) >= 6:
    out_json, key, pattern = sys.argv[3:6]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot, n = 0.0, 0
    for i in order:
        L = launches[i]
        if re.search(pattern, L['kernel']):
            for m, v in L.items():
                if m.startswith('dram__bytes_read.sum') or m.startswith('dram__bytes_write.sum'):
                    tot += v * scale[m.split('[')[1].rstrip(']')]
            n += 1
    d = json.load(open(out_json)) if os.path.exists(out_json) else {}
    d[key] = tot / max(n, 1)
    d[key + '__launches'] = n
    json.dump(d, open(out_json, 'w'), indent=1)
    print(key, '%.1f MB per launch over %d launches' % (d[key] / 1e6, n))
