import sys, json, traceback
sys.path.insert(0, '.')
import torch
import bench
dev = torch.device('cuda', 0)
peaks, _ = bench.measured_peaks()
which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 224
try:
    if which == 'eimp': out = bench.secondary_eimp(dev)
    elif which == 'b1': out = bench.secondary_b1(dev, n)
    elif which == 'pose': out = bench.secondary_pose(dev)
    else: out = bench.secondary_sinkhorn(dev, peaks)
    torch.cuda.synchronize()
    print(json.dumps(out, indent=1))
except Exception:
    traceback.print_exc()
