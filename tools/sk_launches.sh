#!/bin/bash
# per-kernel durations of one Sinkhorn scoring (B=64, N=2000, 20 iters) from an ncu launch list.  usage: sk_launches.sh [fmt]
fmt=${1:-fp24}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/sk_launches_$fmt.csv python tools/sk_one.py $fmt > gpurun_out/sk_launches_$fmt.log 2>&1
python - "$fmt" <<'PY'
import csv, collections, sys
rows = list(csv.reader(open(f'gpurun_out/sk_launches_{sys.argv[1]}.csv')))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    name = d['Kernel Name'].split('(')[0][-40:]
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']; m = d['Metric Name']
    if m == 'gpu__time_duration.sum': v = v / 1e3 if u.startswith('n') else (v if u.startswith('u') else v * 1e3)
    if m.startswith('dram__bytes'): v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}[u]
    a = agg.setdefault(name, collections.defaultdict(list)); a[m].append(v)
for k, a in agg.items():
    n = len(a['gpu__time_duration.sum'])
    avg = lambda m: sum(a[m]) / max(len(a[m]), 1)
    print(f'{k:42s} n={n:3d} avg {avg("gpu__time_duration.sum"):8.1f} us  dram rd {avg("dram__bytes_read.sum"):8.1f} MB wr {avg("dram__bytes_write.sum"):8.1f} MB  issue {avg("smsp__issue_active.avg.pct_of_peak_sustained_active"):5.1f} %')
PY
