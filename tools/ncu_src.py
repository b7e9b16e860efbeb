"""Summarise an ncu --page source --csv dump (SASS view): top instructions by stall samples, with the dominant stall
reason.  usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_src.py [kernel_index] [top_n]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
s = starts[which]
e = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
hdr = rows[s]
print(rows[s - 1][:2] if s > 0 else '')
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
data = []
for r in rows[s + 1:e]:
    if len(r) != len(hdr):
        continue
    try:
        v = float(r[ci['# Samples']] or 0)
    except ValueError:
        continue
    st = sorted(((float(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    data.append((v, r[ci['Source']][:90], st, r[ci['Instructions Executed']]))
tot = sum(d[0] for d in data) or 1
agg = {}
for r in rows[s + 1:e]:
    if len(r) != len(hdr):
        continue
    for c in stall_cols:
        agg[c] = agg.get(c, 0) + float(r[ci[c]] or 0)
print('total samples', tot, ' stall mix:', [(c, round(v / tot, 3)) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]])
for v, src, st, ne in sorted(data, key=lambda x: -x[0])[:topn]:
    print(f'{v:7.0f} {v / tot * 100:5.1f}%  {src:90s} exec={ne:>8s} {[(c, int(x)) for x, c in st]}')
