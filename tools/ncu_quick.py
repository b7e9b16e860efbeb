"""Quick look at one .ncu-rep: key metrics + top stall sites.  usage: python tools/ncu_quick.py <file.ncu-rep> [n_sites]"""
import csv, subprocess, sys
rep = sys.argv[1]
n = sys.argv[2] if len(sys.argv) > 2 else '14'
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(r[idx['Kernel Name']][:90])
    for w in WANT:
        if w in idx:
            print(f'  {w:75s} {r[idx[w]]} {units[idx[w]]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
print(subprocess.run([sys.executable, 'tools/ncu_src.py', '0', n], input=src, capture_output=True, text=True).stdout)
