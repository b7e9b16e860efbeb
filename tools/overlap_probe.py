import sys, torch, time
sys.path.insert(0, '.')
import bench
from imp_release_b200 import DGNNS
from oracle import synth
net = DGNNS(bench.model_config()); net.load_state_dict(synth.make_state_dict('DGNNS', 9, seed=7)); net = net.cuda().eval()
data = {k: v.cuda() for k, v in synth.make_pair_batch(seed=1, batch=64, n0=2000, n1=2000).items()}
def t(n=4):
    for _ in range(3): net(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = net(data)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
with torch.no_grad():
    for ov in (False, True, False, True):
        net.overlap_scoring = ov
        ms, out = t()
        print(f'overlap={ov}: {ms:.2f} ms/step  {64/ms*1e3:.1f} pairs/s  matches {int((out["indices0"][-1] >= 0).sum())}')
    net.overlap_scoring = False; _, a = t(1); net.overlap_scoring = True; _, b = t(1)
    print('same indices:', all(torch.equal(x, y) for x, y in zip(a['indices0'], b['indices0'])))
