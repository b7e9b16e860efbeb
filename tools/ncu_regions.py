"""Stall samples of an ncu --page source dump accumulated over address ranges between marker instructions.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_regions.py   -> prints every instruction with samples
(compact), so that regions can be summed by eye / by grep."""
import csv, sys
rows = list(csv.reader(sys.stdin))
s = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[s]
ci = {h: i for i, h in enumerate(hdr)}
tot = 0
cum = 0
out = []
for r in rows[s + 1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[ci['# Samples']] or 0)
    tot += v
    out.append((r[ci['Address']][-5:], r[ci['Source']].strip()[:70], v, r[ci['Instructions Executed']]))
for a, src, v, ne in out:
    cum += v
    print(f'{a} {v:6.0f} {cum / tot * 100:5.1f}% exec={ne:>8s}  {src}')
