"""GPU: deviation of both attention precision modes from the fp32 oracle in a diffuse (gain 1.0) and a peaky (gain 1.5)
attention regime, plus the cost of the high-precision mode on the bench workload."""
import sys, torch
sys.path.insert(0, '.')
import bench
from imp_release_b200 import DGNNS
from oracle import imp_oracle, synth
for gain in (1.0, 1.5):
    for seed in (21, 33):
        sd = synth.make_state_dict('DGNNS', 9, seed=seed, gain=gain)
        data = synth.make_pair_batch(seed=seed + 1, batch=1, n0=400, n1=380)
        cfg = bench.model_config()
        ref = imp_oracle.Oracle('DGNNS', cfg, sd).forward(data)
        for mode in ('fp16', 'high'):
            net = DGNNS({**cfg, 'attention_precision': mode}); net.load_state_dict(sd); net = net.cuda().eval()
            with torch.no_grad():
                out = net({k: v.cuda() for k, v in data.items()})
            mism = sum(int((a.cpu() != b).sum()) for a, b in zip(out['indices0'], ref['indices0']))
            d = sorted(float((a.cpu() - b).abs().max()) for a, b in zip(out['mscores0'], ref['mscores0']))
            print(f'gain {gain} seed {seed} attention={mode:5s}: index mismatches {mism}, max|dmscore| median {d[4]:.1e} max {d[-1]:.1e}')
sd = synth.make_state_dict('DGNNS', 9, seed=7)
data = {k: v.cuda() for k, v in synth.make_pair_batch(seed=1, batch=64, n0=2000, n1=2000).items()}
for mode in ('fp16', 'high'):
    net = DGNNS({**bench.model_config(), 'attention_precision': mode}); net.load_state_dict(sd); net = net.cuda().eval()
    with torch.no_grad():
        for _ in range(3): net(data)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): net(data)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(f'bench workload attention={mode}: {ms:.1f} ms/step  {64 / ms * 1e3:.0f} pairs/s')
    del net; torch.cuda.empty_cache()
