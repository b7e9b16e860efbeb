"""SuperPoint front-end timing on one B200: whole forward per image (CUDA events) + per-kernel breakdown through the library's
launch spans, next to the reference algorithm on the host cores (oracle port).  usage: python tools/sp_bench.py [H W]"""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
from imp_release_b200 import ops  # noqa: E402
from imp_release_b200.nets.superpoint import SuperPoint  # noqa: E402
from oracle import superpoint_oracle as spo  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 640)
net = SuperPoint({'max_keypoints': 2000})
net.load_state_dict(spo.make_state_dict(11))
net = net.eval().cuda()
img = spo.make_image(21, H, W).cuda()
res = {'image': [H, W]}
with torch.no_grad():
    for _ in range(3):
        out = net({'image': img})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        out = net({'image': img})
    e1.record()
    torch.cuda.synchronize()
    res['forward_ms'] = e0.elapsed_time(e1) / n
    res['keypoints'] = int(out['keypoints'][0].shape[0])
    # dense part only (no host sync inside)
    e0.record()
    for _ in range(n):
        net._dense(img)
    e1.record()
    torch.cuda.synchronize()
    res['dense_ms'] = e0.elapsed_time(e1) / n
    ops.PROFILE = {}
    for _ in range(3):
        net({'image': img})
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
kern = {}
for name, spans in prof.items():
    ms = [a.elapsed_time(b) for a, b, _ in spans]
    work = sum(w or 0 for _, _, w in spans)
    kern[name] = {'calls': len(spans), 'avg_ms': sum(ms) / len(ms), 'total_ms_per_forward': sum(ms) / 3}
    if name.startswith('sp_conv3x3'):
        kern[name]['tflops_algorithmic'] = work / (sum(ms) * 1e9)
res['kernels'] = dict(sorted(kern.items(), key=lambda kv: -kv[1]['total_ms_per_forward']))
conv_flops = sum(w or 0 for k, v in prof.items() if k.startswith('sp_conv3x3') for _, _, w in v) / 3
res['conv3x3_gflop_per_image'] = conv_flops / 1e9
# reference algorithm on the host cores
cfg = {'nms_radius': 4, 'keypoint_threshold': 0.0025, 'remove_borders': 4, 'max_keypoints': 2000}
sd = spo.make_state_dict(11)
cpu_img = img.cpu()
with torch.no_grad():
    spo.forward(sd, cpu_img, cfg)
    t0 = time.time()
    for _ in range(3):
        spo.forward(sd, cpu_img, cfg)
    res['cpu_port_ms'] = (time.time() - t0) / 3 * 1e3
res['cpu_threads'] = torch.get_num_threads()
print(json.dumps(res, indent=1))
json.dump(res, open(f'gpurun_out/sp_bench_{H}x{W}.json', 'w'), indent=1)
