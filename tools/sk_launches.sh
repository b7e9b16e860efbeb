#!/bin/bash
# per-kernel durations of one Sinkhorn call (B=64, N=2000, 20 iters) + the attention probe, from an ncu launch list
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sk_launches.csv python tools/occ_probe.py > gpurun_out/sk_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/sk_launches.csv')))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    name = d['Kernel Name'].split('(')[0][-48:] + ' grid=' + d['Grid Size']
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, v in agg.items(): print(f'{k:75s} n={v[0]:4d} avg {v[1]/v[0]:9.1f} us')
PY
