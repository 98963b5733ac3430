"""Summarise an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline argument uses.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx_summary.csv"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, units, data = rows[0], rows[1], rows[2:]
cols = [H.index('ID'), H.index('Kernel Name')] + [H.index(w) for w in WANT if w in H]
with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow([H[c] for c in cols])
    w.writerow([units[c] for c in cols])
    for r in data:
        w.writerow([r[c][:90] for c in cols])
print(open(out).read())

# optional third argument: JSON file to merge "<key>": average DRAM bytes per launch over the kernels matching a regex
if len(sys.argv) >= 6:
    import json, os, re
    out_json, key, pattern = sys.argv[3], sys.argv[4], sys.argv[5]
    ki, ri, wi = H.index('Kernel Name'), H.index('dram__bytes_read.sum'), H.index('dram__bytes_write.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot, n = 0.0, 0
    for r in data:
        if re.search(pattern, r[ki]):
            tot += float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
            n += 1
    d = json.load(open(out_json)) if os.path.exists(out_json) else {}
    d[key] = tot / max(n, 1)
    d[key + '__launches'] = n
    json.dump(d, open(out_json, 'w'), indent=1)
    print(key, d[key], 'bytes per launch over', n, 'launches')
