"""Turn ncu reports (gpurun_out/<tag>_*.ncu-rep) + launch list into committed summaries under profiles/.
usage: python tools/ncu_summary.py <tag>"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
def to_bytes(v, u):
    m = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(v) * m.get(u, 1)
out = [f'# ncu --set full summaries, tag {tag}\n', 'Captured with `tools/ncu_capture.sh` / `tools/ncu_capture_r02c.sh` (bench workload: 64 pairs, N=2000, 9 iterations; '
       '`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 2 python bench.py --ncu --warmup 1`). '
       'Durations under ncu are serialised / cold-cache; bench.py reports the in-situ CUDA-event times.\n']
traffic = {}
for rep in sorted(f for f in os.listdir('gpurun_out') if f.startswith(tag + '_') and f.endswith('.ncu-rep')):
    raw = subprocess.run(['ncu', '-i', os.path.join('gpurun_out', rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out.append(f'\n## {rep}\n')
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0]
        out.append(f'\n### `{name}` grid {r[idx["launch__grid_size"]]} x block {r[idx["launch__block_size"]]}\n\n| metric | value | unit |\n|---|---:|---|\n')
        for w in WANT:
            if w in idx: out.append(f'| {w} | {r[idx[w]]} | {units[idx[w]]} |\n')
        tb = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        out.append(f'| **dram traffic (read+write)** | {tb/1e6:.1f} | MB |\n')
        fmt = {'0': 'fp32', '1': 'fp16', '2': 'fp24'}
        full = r[idx['Kernel Name']]
        first_arg = full.split('<')[1].split(',')[0].replace('(int)', '').strip() if '<' in full else ''
        key = 'superpoint_conv3x3' if 'conv3x3' in name else 'superpoint_conv1a' if 'conv1a' in name else 'attention' if 'attention_kernel' in name else (('sinkhorn_' + fmt.get(first_arg, 'fp32')) if 'skq_iter' in name or 'sk_ring' in name else ('gemm' if 'gemm' in name else 'instnorm'))
        traffic.setdefault(key, []).append(tb)
    src = subprocess.run(['ncu', '-i', os.path.join('gpurun_out', rep), '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    p = subprocess.run([sys.executable, 'tools/ncu_src.py', '0', '12'], input=src, capture_output=True, text=True).stdout
    out.append('\nTop stall sites (SASS, first captured launch):\n\n```\n' + p + '```\n')
lp = os.path.join('gpurun_out', tag + '_launches.csv')
if os.path.isfile(lp):
    rows = list(csv.reader(open(lp))); hdr = None; agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if 'Kernel Name' in r: hdr = r; continue
        if hdr is None or len(r) != len(hdr): continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum': continue
        v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
        v = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
        a = agg[d['Kernel Name'].split('(')[0][:70]]; a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    out.append(f'\n## launch list ({tag}_launches.csv: `ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900`)\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        out.append(f'| `{k}` | {v[0]} | {v[1]:.3f} | {v[1]/tot:.3f} |\n')
open(f'profiles/{tag}_ncu_summary.md', 'w').write(''.join(out))
# per-launch DRAM traffic for bench.py's roofline.traffic (sinkhorn: one sweep kernel; bench multiplies nothing)
old = json.load(open('profiles/traffic.json')) if os.path.isfile('profiles/traffic.json') else {}
old.update({k: sum(v) / len(v) for k, v in traffic.items()})
json.dump(old, open('profiles/traffic.json', 'w'), indent=1)
print(''.join(out)[:6000])
