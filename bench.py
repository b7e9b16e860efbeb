#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the IMP matching hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

Workload (BASELINE.json configs[1]): DGNNS ("IMP"), 9 iterations, batch of 64 synthetic pairs, N = 2000 keypoints,
D = 256, ``model(data)`` = Sinkhorn + matches at every iteration.  A step = one such forward on one batch.
  value : pairs/s with the inputs already resident in HBM, timed with CUDA events, max over ranks
  e2e   : the same forward through the public API with pinned HOST inputs (H2D inside the timed region) and a
          D2H read of the last iteration's matches (plus, for N > 1, the NCCL gather of all ranks' matches)
Each rank works on its own 64 pairs (weak scaling, replicas + one gather).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_KPTS, N_ITERS, BATCH = 2000, 9, 64
METRIC = 'image-pairs/sec at N=2000 kpts, 9 iters (IMP / DGNNS.forward, Sinkhorn every iteration)'


def model_config(n_layers=N_ITERS):
    return dict(n_layers=n_layers, GNN_layers=['self', 'cross'] * n_layers, norm_fn='in', ac_fn='relu',
                sinkhorn_iterations=20, with_sinkhorn=True, descriptor_dim=256, n_min_tokens=256)


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d, 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        # median of the samples taken under load (upper half: idle samples at the edges are lower)
        under = sm[len(sm) // 2:] if sm else []
        med = under[len(under) // 2] if under else None
        return {'sm_mhz': med, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


def host_cores() -> int:
    """Usable host cores: affinity mask capped by the cgroup CPU quota (os.cpu_count() over-reports in containers)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            q, per = f.read().split()
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return max(1, n)


def attention_flops_per_pair(n, iters):
    from oracle.imp_oracle import attention_flops
    return attention_flops(n, n, iters)


# ----------------------------------------------------------------------------------------------- CPU legs
def cpu_reference_step(n_pairs=1, n=N_KPTS, iters=N_ITERS, repeats=1, seed=1):
    """The reference algorithm (oracle port of nets/gms.py:139-258) on the host cores: seconds per forward."""
    from oracle import imp_oracle, synth
    cfg = model_config(iters)
    sd = synth.make_state_dict('DGNNS', iters, seed=7)
    data = synth.make_pair_batch(seed=seed, batch=n_pairs, n0=n, n1=n)
    orc = imp_oracle.Oracle('DGNNS', cfg, sd)
    best = float('inf')
    with torch.no_grad():
        for _ in range(repeats):
            t = time.perf_counter()
            orc.forward(data)
            best = min(best, time.perf_counter() - t)
    return best


REFERENCE_TIME_CAP_S = 150.0


def run_reference(args, rank, world):
    """The reference algorithm on the host cores.  /root/reference does not exist on the GPU box, so this arm runs the
    oracle port (kind "port": oracle/imp_oracle.py, pinned to the unmodified reference by tests/golden/).  Each step = one
    forward of ONE pair of the 64-pair batch (per-pair CPU time is flat in batch size); --warmup and
    --steps are honoured up to a wall-clock cap so that the run ends within a few minutes on slow hosts."""
    if rank != 0:
        return
    torch.set_num_threads(host_cores())
    cores = torch.get_num_threads()
    t_start = time.perf_counter()
    n_warm = 0
    for _ in range(max(0, args.warmup)):
        cpu_reference_step()
        n_warm += 1
        if time.perf_counter() - t_start > REFERENCE_TIME_CAP_S / 5:
            break
    times = []
    for _ in range(max(1, args.steps)):
        times.append(cpu_reference_step())
        if time.perf_counter() - t_start > REFERENCE_TIME_CAP_S:
            break
    ms = 1e3 * sum(times) / len(times)
    value = 1.0 / (ms / 1e3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': args.gpus,
        'steps': len(times), 'warmup': n_warm, 'steps_requested': args.steps, 'warmup_requested': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'BASELINE.json configs[1]: DGNNS.forward (IMP), N={N_KPTS}, D=256, {N_ITERS} iters, Sinkhorn(20)+matches '
                               f'every iteration; each step = a bounded sample of 1 pair of the {BATCH}-pair batch (per-pair CPU '
                               f'time is flat in batch size, BASELINE.md section 3)'},
        'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{len(times)} steps of 1 pair each (mean), oracle/imp_oracle.py (torch CPU fp32, all host '
                                   f'threads); time cap {REFERENCE_TIME_CAP_S:.0f} s'},
        'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- secondary configs
def _timed(fn, warm, n):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_secondary(dev, peaks, n_pairs=224):
    """BASELINE.json configs[2..4] on one GPU (device-resident inputs, CUDA-event timed, >= 2 warm-ups).  configs[3] is the
    one-pair-per-call evaluation shape of eval/eval_imp.py:155-173 on ONE rank: >= 200 pairs with distinct ragged keypoint
    counts, 15 iterations, produce_matches(only_last=True)."""
    out = {}

    def leg(name, fn):       # a failing secondary configuration must never take the headline line down with it
        try:
            out.update(fn())
        except Exception as e:  # noqa: BLE001
            out[name] = {'error': f'{type(e).__name__}: {e}'[:300]}
            print(f'[secondary] {name} failed: {e}', file=sys.stderr, flush=True)
        try:
            torch.cuda.empty_cache()
        except Exception:  # noqa: BLE001  (a sticky CUDA error: the headline numbers are already measured)
            pass

    leg('configs[2] EIMP', lambda: secondary_eimp(dev))
    leg('configs[3] one rank', lambda: secondary_b1(dev, n_pairs))
    leg('configs[4] Sinkhorn-only', lambda: secondary_sinkhorn(dev, peaks))
    leg('eval loop with host pose', lambda: secondary_pose(dev))
    leg('SuperPoint front-end', lambda: secondary_superpoint(dev))
    return out


def secondary_eimp(dev):
    from imp_release_b200 import AdaGMN
    from oracle import synth
    out = {}
    # ---- configs[2]: EIMP, N = 2000 -> pruned, 9 iterations, batch 128
    B = 128
    net = AdaGMN(model_config(9))
    net.load_state_dict(synth.make_state_dict('AdaGMN', 9, seed=7, bin_score=8.0))
    net = net.to(dev).eval()
    data = {k: v.to(dev) for k, v in synth.make_pair_batch(seed=2, batch=B, n0=N_KPTS, n1=N_KPTS).items()}
    with torch.no_grad():
        ms = _timed(lambda: net(data), 2, 3)
    cnt, _ = net._kept
    out['configs[2] EIMP batch=128 N=2000 9 iters'] = {
        'pairs_per_s': B / ms * 1e3, 'ms_per_batch': ms, 'kept_keypoints_mean': float(cnt.float().mean()),
        'kept_min': int(cnt.min()), 'kept_max': int(cnt.max()), 'bin_score': 8.0,
        'note': 'AdaGMN.forward, bin_score raised to 8 so that the seeded random weights prune (SURVEY.md 8(d))'}
    return out


def secondary_b1(dev, n_pairs=224):
    from imp_release_b200 import DGNNS
    from imp_release_b200.graphed import LatencyMatcher
    from oracle import synth
    out = {}
    # ---- configs[3] on one rank: DGNNS 15 iterations, one pair per call, ragged N in [1200, 2000]
    net = DGNNS(model_config(15))
    net.load_state_dict(synth.make_state_dict('DGNNS', 15, seed=7))
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(11)
    sizes = [(int(a), int(b)) for a, b in torch.randint(1200, 2001, (n_pairs, 2), generator=g).tolist()]
    big = {k: v.to(dev) for k, v in synth.make_pair_batch(seed=5, batch=1, n0=2000, n1=2000, width=1600, height=1200).items()}

    def pair(i):
        n0, n1 = sizes[i]
        return {'descriptors0': big['descriptors0'][:, :n0], 'descriptors1': big['descriptors1'][:, :n1],
                'keypoints0': big['keypoints0'][:, :n0], 'keypoints1': big['keypoints1'][:, :n1],
                'scores0': big['scores0'][:, :n0], 'scores1': big['scores1'][:, :n1], 'image0': big['image0'], 'image1': big['image1']}
    pairs = [pair(i) for i in range(n_pairs)]

    def sweep(fn, tag=''):
        if os.environ.get('IMP_BENCH_DEBUG'):
            torch.cuda.synchronize()
            print('[secondary_b1] sweep', tag, file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = [fn(d) for d in pairs]
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n_pairs, (time.perf_counter() - t0) * 1e3 / n_pairs, res
    with torch.no_grad():
        eager = lambda d: net.produce_matches(d, p=0.2, only_last=True)['indices0'][-1]
        sweep(eager, 'eager warm')                               # warm-up: workspaces of every bucket
        ms_eager, wall_eager, ref = sweep(eager, 'eager')
        lm1 = LatencyMatcher(net, slots=1)
        t0 = time.perf_counter()
        sweep(lambda d: lm1(d)['indices0'][-1], 'graph1 build')  # builds the graphs (one per 128-keypoint bucket)
        build_s = time.perf_counter() - t0
        ms_g1, wall_g1, r1 = sweep(lambda d: lm1(d)['indices0'][-1], 'graph1')
        multi, host_ms = {}, {}
        same4 = True
        for n_slots in (4, 8):
            lm4 = LatencyMatcher(net, slots=n_slots)
            sweep(lambda d: lm4.submit(d), f'graph{n_slots} build')
            torch.cuda.synchronize()
            t_h = time.perf_counter()
            for d in pairs[:64]:                                   # host cost of a submit (no synchronisation inside)
                lm4.submit(d)
            host_ms[n_slots] = (time.perf_counter() - t_h) * 1e3 / 64
            _, wall_g4, tickets = sweep(lambda d: lm4.submit(d), f'graph{n_slots}')
            r4 = [lm4.result(t)['indices0'][-1] for t in tickets]
            torch.cuda.synchronize()
            same4 = same4 and all(torch.equal(a, b) for a, b in zip(ref, r4))
            multi[n_slots] = (wall_g4, lm4.captures)
            del lm4, tickets, r4
        same1 = all(torch.equal(a, b) for a, b in zip(ref, r1))
        best_slots = min(multi, key=lambda k: multi[k][0])
    out['configs[3] one rank: IMP 15 iters, 1 pair per call, ragged N0,N1 in [1200,2000]'] = {
        'pairs': n_pairs, 'distinct_shapes': len(set(sizes)),
        'eager_ms_per_pair': ms_eager, 'eager_wall_ms_per_pair': wall_eager,
        'graph_1_slot_ms_per_pair': ms_g1, 'graph_1_slot_wall_ms_per_pair': wall_g1,
        'graph_4_slots_wall_ms_per_pair': multi[4][0], 'graph_8_slots_wall_ms_per_pair': multi[8][0],
        'pairs_per_s_in_flight': 1e3 / multi[best_slots][0], 'best_slots': best_slots,
        'host_ms_per_submit': host_ms[best_slots],
        'graphs_captured_1_slot': lm1.captures, 'graphs_captured_4_slots': multi[4][1], 'first_sweep_incl_capture_s': build_s,
        'graph_results_equal_eager': bool(same1 and same4),
        'note': 'DGNNS.produce_matches(only_last=True); LatencyMatcher = bucketed (128) static shapes + CUDA-graph replay, '
                'k slots = k pairs in flight on k streams (wall clock incl. the final synchronize: CUDA events on one stream do not see the '
                'others); device-resident inputs; ms = CUDA events, wall = host clock'}
    return out


def secondary_pose(dev, n_pairs=32):
    """SURVEY.md 8(f) rank 1, measured: the one-pair-per-call evaluation loop INCLUDING the host-side pose step
    (cv2 essential-matrix RANSAC + cheirality, imp_release_b200/host_pose.py = the reference's eval/pose_estimation.py:92-115)
    on synthetic two-view scenes.  serial = what eval/eval_imp.py:155-173 does (match, blocking copy, RANSAC, next pair);
    overlapped = LatencyMatcher (4 pairs in flight) + PoseOverlap (RANSAC in worker threads on async-copied matches)."""
    import cv2  # noqa: F401
    from imp_release_b200 import DGNNS, host_pose
    from imp_release_b200.graphed import LatencyMatcher
    from oracle import synth
    net = DGNNS(model_config(15))
    net.load_state_dict(synth.make_state_dict('DGNNS', 15, seed=7))
    net = net.to(dev).eval()
    scenes = [synth.make_scene_pair(300 + i, 1400 + 17 * (i % 20), 1500 - 13 * (i % 20)) for i in range(n_pairs)]
    feed = [{k: (v.to(dev) if torch.is_tensor(v) and k != 'perm' and not k.startswith('image') else v) for k, v in sc.items()}
            for sc in scenes]
    workers = min(8, host_cores())
    with torch.no_grad():
        for d in feed[:3]:
            net.produce_matches(d, p=0.2, only_last=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_match = 0
        for d in feed:
            o = net.produce_matches(d, p=0.2, only_last=True)
            i0 = o['indices0'][-1][0].cpu().numpy()
            n_match += int((i0 > -1).sum())
            host_pose.pose_from_matches(i0, None, d['pts0_cpu'], d['pts1_cpu'], d['K0'], d['K1'])
        serial = time.perf_counter() - t0
        lm = LatencyMatcher(net, slots=4)
        host_pose.evaluate_pairs(lm, feed, workers=workers)            # builds the graphs
        t0 = time.perf_counter()
        res = host_pose.evaluate_pairs(lm, feed, workers=workers)
        overlapped = time.perf_counter() - t0
    return {'eval loop with host pose (SURVEY 8(f) rank 1): IMP 15 iters + cv2 RANSAC per pair': {
        'pairs': n_pairs, 'matches_per_pair_mean': n_match / n_pairs, 'poses_found': sum(r is not None for r in res),
        'serial_pairs_per_s': n_pairs / serial, 'overlapped_pairs_per_s': n_pairs / overlapped, 'pose_workers': workers,
        'note': 'serial = match -> blocking D2H -> RANSAC -> next pair (the reference loop); overlapped = 4 pairs in flight on '
                'the GPU, RANSAC in worker threads; wall clock, device-resident inputs'}}


def secondary_superpoint(dev):
    """SURVEY.md 8(f) rank 2, measured: the SuperPoint front-end (nets/superpoint.py) and the whole image pair -> matches
    pipeline on the GPU (SuperPoint on both images, IMP 15 iterations, produce_matches(only_last=True)); the reference algorithm
    (oracle/superpoint_oracle.py, a port of the reference class pinned by its goldens) timed on the host cores beside it."""
    from imp_release_b200 import DGNNS
    from imp_release_b200.nets.superpoint import SuperPoint
    from oracle import superpoint_oracle as spo
    from oracle import synth
    out = {}
    sp = SuperPoint({'max_keypoints': N_KPTS})
    sp.load_state_dict(spo.make_state_dict(11))
    sp = sp.to(dev).eval()
    with torch.no_grad():
        for (H, W) in ((480, 640), (1200, 1600)):
            img = spo.make_image(21, H, W).to(dev)
            ms = _timed(lambda: sp({'image': img}), 3, 10)
            dense = _timed(lambda: sp._dense(img), 3, 10)
            row = {'forward_ms': ms, 'dense_part_ms': dense, 'keypoints': int(sp({'image': img})['keypoints'][0].shape[0])}
            if H == 480:
                sd, cfg, cpu_img = spo.make_state_dict(11), {'nms_radius': 4, 'keypoint_threshold': 0.0025, 'remove_borders': 4,
                                                            'max_keypoints': N_KPTS}, img.cpu()
                spo.forward(sd, cpu_img, cfg)
                t0 = time.perf_counter()
                spo.forward(sd, cpu_img, cfg)
                row['cpu_port_ms'] = (time.perf_counter() - t0) * 1e3
                row['cpu_threads'] = torch.get_num_threads()
            out[f'SuperPoint front-end (SURVEY 8(f) rank 2), {H}x{W} image, top-{N_KPTS} keypoints'] = row
        # image pair -> matches, everything on the GPU
        net = DGNNS(model_config(15))
        net.load_state_dict(synth.make_state_dict('DGNNS', 15, seed=7))
        net = net.to(dev).eval()
        imgs = [spo.make_image(30 + i, 1200, 1600).to(dev) for i in range(2)]
        shape = torch.zeros(1, 1, 1200, 1600)

        def two_step():          # the reference's structure: extract (host reads the keypoint counts), then match
            f = [sp({'image': im}) for im in imgs]
            data = {'image0': shape, 'image1': shape}
            for i in (0, 1):
                data[f'keypoints{i}'] = f[i]['keypoints'][0][None]
                data[f'scores{i}'] = f[i]['scores'][0][None]
                data[f'descriptors{i}'] = f[i]['descriptors'][0].t()[None].contiguous()
            return net.produce_matches(data, p=0.2, only_last=True)

        from imp_release_b200.pipeline import ImagePairMatcher
        ipm = ImagePairMatcher(sp, net)
        ms_two = _timed(two_step, 3, 10)
        ms_pipe = _timed(lambda: ipm(imgs[0], imgs[1]), 3, 10)
        # several pairs in flight: one stream + one matcher replica (own workspaces, shared weights) per slot, one host thread
        slots = 4
        streams = [torch.cuda.Stream(dev) for _ in range(slots)]
        ipms = ImagePairMatcher.slots(sp, net, slots)
        n_pairs = 32

        def in_flight():
            res = []
            for i in range(n_pairs):
                with torch.cuda.stream(streams[i % slots]):
                    res.append(ipms[i % slots](imgs[0], imgs[1]))
            return res

        in_flight()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = in_flight()
        torch.cuda.synchronize()
        ms_flight = (time.perf_counter() - t0) * 1e3 / n_pairs
        same = all(torch.equal(r['indices0'], res[0]['indices0']) for r in res[1:])
        # ... and as CUDA-graph replay (one graph per slot holds both detections and the matcher: ~600 launches per pair)
        from imp_release_b200.pipeline import GraphedImagePairMatcher
        gm = GraphedImagePairMatcher(sp, net, slots=slots)
        for _ in range(2 * slots):
            gm(imgs[0], imgs[1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tickets = [gm.submit(imgs[0], imgs[1]) for _ in range(n_pairs)]
        gres = [gm.result(t) for t in tickets]
        torch.cuda.synchronize()
        ms_graph = (time.perf_counter() - t0) * 1e3 / n_pairs
        same = same and all(torch.equal(r['indices0'], res[0]['indices0']) for r in gres)
        out['image pair -> matches on the GPU (2 x SuperPoint 1200x1600 + IMP 15 iters, one pair per call, eager)'] = {
            'two_step_ms_per_pair': ms_two, 'no_host_sync_ms_per_pair': ms_pipe, 'no_host_sync_4_in_flight_ms_per_pair': ms_flight,
            'graph_replay_4_in_flight_ms_per_pair': ms_graph, 'in_flight_results_identical': same, 'keypoints_per_image': N_KPTS,
            'note': 'two_step = SuperPoint.forward (host reads the keypoint counts) then produce_matches; no_host_sync = '
                    'imp_release_b200.pipeline.ImagePairMatcher (device-side counts, the host only enqueues): 10 pairs '
                    'back to back between two CUDA events; 4_in_flight = 32 pairs round-robin over 4 streams, wall clock; graph_replay = '
                    'the same with one CUDA graph per slot (pipeline.GraphedImagePairMatcher)'}
    return out


def secondary_sinkhorn(dev, peaks):
    from imp_release_b200 import ops
    out = {}
    # ---- configs[4]: Sinkhorn only, 2048 x 2048 padded (dist 2047^2), 100 iterations
    for Bs in (16, 1):
        N, ld = 2047, 2048
        dist_ = torch.randn(Bs, N, ld, device=dev, generator=torch.Generator(dev).manual_seed(3)) * 3
        ws = ops.SinkhornWorkspace(Bs, N, N, dev)
        bs = torch.tensor(1.0, device=dev)
        ms = _timed(lambda: ops.sinkhorn(dist_, ld, bs, 100, ws, write_scores=True), 2, 3)
        mat = 4.0 * Bs * 2048 * 2048
        streaming = ws.q_store is not None
        sweeps = 101 if streaming else None          # init + 99 iteration sweeps + final (one sweep per iteration)
        out[f'configs[4] Sinkhorn-only 2048^2 x 100 iters, batch={Bs}'] = {
            'ms': ms, 'path': 'streaming kernels, fp32 copy (matrix > L2)' if streaming else 'shared-memory resident, one cooperative launch',
            'bytes_moved_GBps': (sweeps + 2) * mat / ms / 1e6 if streaming else mat * 2 / ms / 1e6,
            'frac_of_hbm_peak': ((sweeps + 2) * mat / ms / 1e6 / peaks['hbm_gbs']) if streaming else None,
            'reference_counting_GBps': 2 * 100 * mat / ms / 1e6,
            'note': 'bytes_moved = one 4-byte read of the matrix per sweep (+ the init write and the final score write); '
                    'reference_counting = 2 sweeps per iteration as SURVEY.md 8(d) counts the reference formulation'
                    if streaming else 'the 16.8 MB matrix is read from HBM once and lives in shared memory for all 100 iterations: '
                                      'grid-barrier latency bound, HBM traffic is 2 passes'}
        del ws, dist_
    return out


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, world, local_rank):
    import torch.distributed as dist
    from imp_release_b200 import DGNNS, ops
    from oracle import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sk_storage = args.sinkhorn_storage or ops.default_sk_storage()
    cfg = {**model_config(), 'attention_precision': args.attention_precision, 'sinkhorn_storage': sk_storage}
    net = DGNNS(cfg)
    net.load_state_dict(synth.make_state_dict('DGNNS', N_ITERS, seed=7), strict=True)
    net = net.to(dev).eval()

    host = synth.make_pair_batch(seed=1 + rank, batch=BATCH, n0=N_KPTS, n1=N_KPTS)
    keys = ['descriptors0', 'descriptors1', 'keypoints0', 'keypoints1', 'scores0', 'scores1']
    pinned = {k: host[k].pin_memory() for k in keys}
    shapes = {'image0': host['image0'], 'image1': host['image1']}      # only .shape is read (nets/gms.py:162-164)
    resident = {k: v.to(dev) for k, v in pinned.items()}
    resident.update(shapes)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    n_pairs_total = BATCH * world

    def step_resident():
        with torch.no_grad():
            return net(resident)

    out_i = torch.empty(BATCH, N_KPTS, dtype=torch.int64).pin_memory()
    out_s = torch.empty(BATCH, N_KPTS, dtype=torch.float32).pin_memory()

    # End to end = what a data loader in front of the public API does: every step stages one full batch of pinned host
    # tensors to the device (PairFeeder: double-buffered on a copy stream, so the H2D of step i+1 overlaps the matcher of
    # step i) and reads the step's matches back into pinned host memory.  One H2D and one D2H per step, all inside the timed
    # region; `e2e.serial` below is the same without the overlap (blocking copies on the compute stream).
    from imp_release_b200.feeder import PairFeeder
    feeder = PairFeeder(dev, depth=2)
    host_batch = dict(pinned)
    host_batch.update(shapes)

    def finish(out):
        i0, s0 = out['indices0'][-1], out['mscores0'][-1]
        if world > 1:
            # fixed-stride gather of every rank's matches to rank 0 (24 KB per pair), cf. shard.gather_matches
            gi = [torch.empty_like(i0) for _ in range(world)] if rank == 0 else None
            gs = [torch.empty_like(s0) for _ in range(world)] if rank == 0 else None
            dist.gather(i0, gi, dst=0)
            dist.gather(s0, gs, dst=0)
        out_i.copy_(i0, non_blocking=True)
        out_s.copy_(s0, non_blocking=True)

    def step_e2e():
        with torch.no_grad():
            d = feeder.next()              # staged during the previous step (or by the priming call below)
            feeder.stage(host_batch)       # H2D of the next step's inputs, overlapping this step's kernels
            out = net(d)
            finish(out)
        return out

    def step_e2e_serial():
        with torch.no_grad():
            d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            d.update(shapes)
            out = net(d)
            finish(out)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, wall

    for _ in range(args.warmup):
        step_resident()
    if args.ncu:
        step_resident()
        torch.cuda.synchronize()
        return
    feeder.stage(host_batch)         # prime the pipeline: from here on every step_e2e() issues exactly one H2D batch
    step_e2e()
    step_e2e_serial()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.LAUNCHES
    ms_dev, _ = timed(step_resident, args.steps)
    launches = ops.LAUNCHES - launches0
    clocks = sampler.stop()
    ms_e2e, _ = timed(step_e2e, args.steps)
    ms_e2e_serial, _ = timed(step_e2e_serial, args.steps)

    # per-kernel breakdown with CUDA events on the launching stream (separate pass so `value` is unperturbed)
    ops.PROFILE = {}
    net.overlap_scoring = False      # serialise the two streams so that per-kernel event times are uncontended
    from imp_release_b200 import _lib as implib
    implib.load().imp_set_profiling(1)
    barrier()
    for _ in range(max(1, min(args.steps, 3))):
        step_resident()
    torch.cuda.synchronize()
    prof = {}
    for name, spans in ops.PROFILE.items():
        tot = sum(a.elapsed_time(b) for a, b, _ in spans)
        work = sum(w for _, _, w in spans)
        prof[name] = (tot, len(spans), work)
    ops.PROFILE = None
    net.overlap_scoring = True
    sk_iter_ms = float(implib.load().imp_sinkhorn_iter_ms())
    implib.load().imp_set_profiling(0)
    total_prof = sum(v[0] for v in prof.values()) or 1.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    value = n_pairs_total / (ms_dev / 1e3)
    e2e_value = n_pairs_total / (ms_e2e / 1e3)
    top = max(prof.items(), key=lambda kv: kv[1][0])[0]
    kernels = {k: {'ms_per_step_share': round(v[0] / total_prof, 4), 'avg_ms': round(v[0] / v[1], 4), 'calls': v[1]}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}

    def roof(name):
        tot, n, work = prof[name]
        avg_s = tot / n / 1e3
        if name.startswith('sinkhorn'):
            # dominant kernel of the group: skq_iter_kernel, one launch per Sinkhorn iteration = ONE sweep over the stored
            # copy of softmax(M) [B, N0+1, roundup16(N1+1)].  `achieved` = bytes that sweep moves / its CUDA-event duration.
            # With the default fp32 copy this equals the algorithmic figure of SURVEY.md 8(d) per sweep (4 B x (N0+1)(N1+1))
            # up to 0.7 % of row padding.
            sk_bytes = {'fp32': 4, 'fp24': 3, 'fp16': 2}[sk_storage]
            ldq = (N_KPTS + 1 + 15) // 16 * 16
            moved = float(sk_bytes) * BATCH * (N_KPTS + 1) * ldq
            algo = 4.0 * BATCH * (N_KPTS + 1) * (N_KPTS + 1)
            n_launch = 21           # init + 19 iterations + final (the column arg-max is fused into the final pass)
            it_s = (sk_iter_ms if sk_iter_ms > 0 else (tot / n) / n_launch) / 1e3
            ach = moved / it_s / 1e9
            return {'kernel': f'skq_iter_kernel<{sk_storage}> (one Sinkhorn iteration, 19 of the {n_launch} launches of a scoring)',
                    'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                    'frac': ach / peaks['hbm_gbs'], 'traffic': None, 'launch_ms': it_s * 1e3,
                    'bytes_per_launch': moved, 'algorithmic_bytes_per_launch': algo,
                    'storage': {'format': sk_storage, 'bytes_per_element': sk_bytes},
                    'reference_counting': {'GBps': 2 * algo / it_s / 1e9,
                                           'note': 'side note only: the reference formulation (SURVEY.md 8(d)) sweeps the matrix '
                                                   'twice per iteration (row pass + column pass); this kernel fuses both into one'},
                    'whole_scoring': {'ms': tot / n, 'launches': n_launch},
                    'note': 'achieved = bytes of the stored copy one sweep reads (%d B x B x (N0+1) x roundup16(N1+1)) / mean '
                            'CUDA-event duration of the 19 iteration launches of each scoring (events recorded inside '
                            'libimp_b200.so on the launching stream); peak = STREAM-style copy of ' % sk_bytes + peak_src}
        pk = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops'])
        ach = work / n / avg_s / 1e12
        return {'kernel': name, 'bound': 'tensor', 'achieved': ach, 'peak': pk, 'unit': 'TFLOP/s', 'frac': ach / pk,
                'traffic': None, 'note': 'algorithmic FLOPs / launch (SURVEY.md 8(d)) over the CUDA-event launch time; '
                                         'peak = sustained dense bf16/fp16 cuBLAS of ' + peak_src}
    roofline = roof(top)
    extra = {k: roof(k) for k in ('attention', 'attention_shared', 'sinkhorn') if k in prof and k != top}
    tr = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.isfile(tr):
        with open(tr) as f:
            t = json.load(f)
        for r in [roofline] + list(extra.values()):
            key = ('sinkhorn_' + sk_storage) if r['kernel'].startswith('sk') else r['kernel']
            if key in t:
                r['traffic'] = t[key]
    secondary = None
    if world == 1 and not args.no_secondary:
        del resident, feeder
        net._engine = None
        torch.cuda.empty_cache()
        secondary = run_secondary(dev, peaks)

    torch.set_num_threads(host_cores())
    cpu_s = cpu_reference_step(repeats=2) if world == 1 and not args.no_cpu_baseline else None
    cpu_baseline = None
    if cpu_s is not None:
        cpu_baseline = {'value': 1.0 / cpu_s, 'unit': 'pairs/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                        'sample': 'best of 2 forwards of 1 pair (N=2000, 9 iters) with oracle/imp_oracle.py on all host '
                                  'threads; per-pair CPU time is flat in batch size'}
    att = attention_flops_per_pair(N_KPTS, N_ITERS)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16x3 split (fp32-equivalent) projections/scores, '
                 + ('f16x3 split' if args.attention_precision == 'high' else 'f16') +
                 ' attention operands, f32 accumulate/softmax/Sinkhorn (Sinkhorn sweeps read a ' + sk_storage + ' copy of softmax(M))',
        'data': 'synthetic',
        'config': {'workload': f'BASELINE.json configs[1]: DGNNS.forward (IMP), batch={BATCH} pairs/GPU, N={N_KPTS}, D=256, '
                               f'{N_ITERS} iters, Sinkhorn(20)+matches every iteration',
                   'l2': 'per-step inputs (264 MB) and working set (>5 GB) exceed the 126 MB L2; no explicit flush',
                   'parallelism': f'replicas x{world}, pairs sharded by rank, one gather of matches'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': out_i.numel() * 8 + out_s.numel() * 4,
                'pipelining': 'PairFeeder: the H2D of step i+1 (pinned -> device, copy stream) overlaps the matcher of step i; '
                              'one full-batch H2D and one D2H of the matches per timed step',
                'serial': {'value': n_pairs_total / (ms_e2e_serial / 1e3), 'ms_per_step': ms_e2e_serial,
                           'note': 'blocking copies on the compute stream, as the reference feeds its model'}},
        'gpu_launches': launches,
        'roofline': roofline,
        'roofline_other': extra,
        'attention_tflops_whole_step': att * n_pairs_total / (ms_dev / 1e3) / 1e12,
        'kernels': kernels,
        'cpu_baseline': cpu_baseline,
        'secondary': secondary,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The driver parses ONE JSON line from stdout: everything else (NCCL banners, warnings) goes to stderr."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)            # library chatter printed to fd 1 (e.g. "NCCL version ...") lands on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip BASELINE.json configs[2..4] (run by default at N=1)')
    ap.add_argument('--attention-precision', default='fp16', choices=['fp16', 'high'],
                    help="'high' = split-precision attention (DESIGN.md section 2); default = the fast fp16 mode")
    ap.add_argument('--sinkhorn-storage', default=None, choices=['fp32', 'fp24', 'fp16'],
                    help='storage of softmax(M) for the Sinkhorn iteration sweeps (DESIGN.md section 2); default = library default')
    ap.add_argument('--ncu', action='store_true', help='profiling aid: 1 warm-up + 1 resident step, nothing else (numbers invalid)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    if args.warmup < 3 and not args.ncu:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)')
    run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
