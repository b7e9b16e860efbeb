"""End-to-end parity (GPU): GM / DGNNS / AdaGMN through the reference-shaped Python API vs the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from imp_release_b200 import GM, DGNNS, AdaGMN  # noqa: E402
from oracle import imp_oracle, synth  # noqa: E402


def cfg(nl, **kw):
    c = dict(n_layers=nl, GNN_layers=['self', 'cross'] * nl, norm_fn='in', ac_fn='relu', sinkhorn_iterations=20,
             with_sinkhorn=True, descriptor_dim=256)
    c.update(kw)
    return c


def to_cuda(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def compare(out, ref, tol=1e-3):
    assert len(out['indices0']) == len(ref['indices0'])
    for ni, (a, b) in enumerate(zip(out['indices0'], ref['indices0'])):
        assert a.dtype == torch.int64
        mism = int((a.cpu() != b).sum())
        assert mism == 0, f'iteration {ni}: {mism} index mismatches'
    for ni, (a, b) in enumerate(zip(out['mscores0'], ref['mscores0'])):
        d = float((a.cpu() - b).abs().max())
        assert d < tol, f'iteration {ni}: mscores differ by {d}'


@pytest.mark.parametrize('N0,N1,B', [(512, 512, 1)])
def test_gm_config1(N0, N1, B, sk_path):
    """BASELINE.json configs[0]: GM.forward, N=512, 3 iterations."""
    c = cfg(3)
    sd = synth.make_state_dict('GM', 3, seed=5)
    data = synth.make_pair_batch(seed=0, batch=B, n0=N0, n1=N1)
    ref = imp_oracle.Oracle('GM', c, sd).forward(data)
    net = GM(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net(to_cuda(data))
    compare(out, ref)
    for a, b in zip(out['scores'], ref['scores']):
        assert a.shape == b.shape
        assert float((a.cpu()[:, :-1, :-1] - b[:, :-1, :-1]).abs().max()) < 1e-3


@pytest.mark.parametrize('N0,N1,B,nl,only_last', [(500, 460, 2, 9, False), (300, 260, 1, 9, True), (1000, 960, 1, 9, False),
                                                   (700, 900, 1, 15, True)])
def test_dgnns(N0, N1, B, nl, only_last, sk_path):
    c = cfg(nl)
    sd = synth.make_state_dict('DGNNS', nl, seed=7)
    data = synth.make_pair_batch(seed=3, batch=B, n0=N0, n1=N1)
    ref = imp_oracle.Oracle('DGNNS', c, sd).produce_matches(data, p=0.2, only_last=only_last)
    net = DGNNS(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net.produce_matches(to_cuda(data), p=0.2, only_last=only_last)
    compare(out, ref)


@pytest.mark.parametrize('N0,N1,B,seed', [(500, 460, 2, 7), (1000, 960, 1, 11)])
def test_adagmn_batched_pruning(N0, N1, B, seed, sk_path):
    """BASELINE.json configs[2] shape: EIMP with adaptive pooling (bin_score raised so that pruning happens)."""
    c = cfg(9, n_min_tokens=256)
    sd = synth.make_state_dict('AdaGMN', 9, seed=seed, bin_score=6.0)
    data = synth.make_pair_batch(seed=seed + 1, batch=B, n0=N0, n1=N1)
    ref = imp_oracle.Oracle('AdaGMN', c, sd).forward(data)
    net = AdaGMN(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net(to_cuda(data))
    cnt, ids = net._kept
    kept = cnt.cpu().tolist()
    assert kept[:B] == ref['kept0'][-1] and kept[B:] == ref['kept1'][-1], (kept, ref['kept0'][-1], ref['kept1'][-1])
    assert min(kept) < min(N0, N1), 'test vector must actually prune'
    compare(out, ref)
    assert out['scores'][0].shape == ref['scores'][0].shape


def test_cuda_graph_replay_matches_eager():
    from imp_release_b200.graphed import GraphedMatcher
    c = cfg(9)
    sd = synth.make_state_dict('DGNNS', 9, seed=7)
    net = DGNNS(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    d1 = to_cuda(synth.make_pair_batch(seed=31, batch=1, n0=700, n1=640))
    d2 = to_cuda(synth.make_pair_batch(seed=32, batch=1, n0=700, n1=640))
    g = GraphedMatcher(net, d1, p=0.2, only_last=True)
    for d in (d2, d1, d2):
        with torch.no_grad():
            eager = net.produce_matches(d, p=0.2, only_last=True)
            i_e, m_e = eager['indices0'][-1].clone(), eager['mscores0'][-1].clone()
        out = g(d)
        torch.cuda.synchronize()
        assert torch.equal(out['indices0'][-1], i_e)
        assert float((out['mscores0'][-1] - m_e).abs().max()) < 1e-5
    ref = imp_oracle.Oracle('DGNNS', c, sd).produce_matches({k: v.cpu() for k, v in d2.items()}, only_last=True)
    assert torch.equal(out['indices0'][-1].cpu(), ref['indices0'][-1])


@pytest.mark.parametrize('N0,N1,B,nl', [(5, 9, 1, 3), (129, 64, 3, 4), (2300, 2210, 1, 3)])
def test_dgnns_edge_sizes(N0, N1, B, nl, sk_path):
    """Tiny, odd-batch and > 2048-keypoint pairs (longest Sinkhorn row variant, > 16 key tiles)."""
    c = cfg(nl)
    sd = synth.make_state_dict('DGNNS', nl, seed=13)
    data = synth.make_pair_batch(seed=17, batch=B, n0=N0, n1=N1)
    ref = imp_oracle.Oracle('DGNNS', c, sd).produce_matches(data, p=0.2, only_last=False)
    net = DGNNS(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net.produce_matches(to_cuda(data), p=0.2, only_last=False)
    compare(out, ref)


def test_dual_softmax_model_path():
    """with_sinkhorn=False (eval_imp.py --use_dual_softmax): dual-softmax scorer through the model API."""
    c = cfg(3, with_sinkhorn=False)
    sd = synth.make_state_dict('DGNNS', 3, seed=5)
    data = synth.make_pair_batch(seed=6, batch=2, n0=150, n1=170)
    ref = imp_oracle.Oracle('DGNNS', c, sd).produce_matches(data, p=0.2)
    net = DGNNS(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net.produce_matches(to_cuda(data), p=0.2)
    compare(out, ref)


def test_run_adapters_and_weight_reload():
    """mode=1 `run` adapters (eval/eval_yfcc_full.py:54) and repacking after load_state_dict."""
    c = cfg(3)
    data = synth.make_pair_batch(seed=9, batch=1, n0=200, n1=180)
    nk0 = imp_oracle.normalize_keypoints(data['keypoints0'], data['image0'].shape)
    nk1 = imp_oracle.normalize_keypoints(data['keypoints1'], data['image1'].shape)
    rd = {'desc1': data['descriptors0'], 'desc2': data['descriptors1'],
          'x1': torch.cat([nk0, data['scores0'][..., None]], -1), 'x2': torch.cat([nk1, data['scores1'][..., None]], -1)}
    net = DGNNS(c).cuda().eval()
    for seed in (3, 4):                      # second pass: new weights must be repacked
        sd = synth.make_state_dict('DGNNS', 3, seed=seed)
        net.load_state_dict(sd, strict=True)
        ref = imp_oracle.Oracle('DGNNS', c, sd).produce_matches(
            {**data, 'norm_keypoints0': nk0, 'norm_keypoints1': nk1}, p=0.2, only_last=True)
        with torch.no_grad():
            out = net(to_cuda(rd), mode=1)
        ri = ref['indices0'][-1][0]
        idx0 = torch.where(ri >= 0)[0]
        assert torch.equal(out['index0'].cpu(), idx0) and torch.equal(out['index1'].cpu(), ri[idx0])


@pytest.mark.parametrize('kind', ['DGNNS', 'AdaGMN'])
def test_high_precision_attention_with_peaky_weights(kind):
    """Weights scaled x1.5 make the attention sharply peaked (entropy ~2 nats instead of ~5.6): the single-fp16
    attention drifts to ~1e-3 in the scores there, the 'high' mode stays at fp32 level."""
    c = cfg(9, attention_precision='high')
    sd = synth.make_state_dict(kind, 9, seed=21, gain=1.5, bin_score=(4.0 if kind == 'AdaGMN' else 1.0))
    data = synth.make_pair_batch(seed=22, batch=1, n0=400, n1=380)
    ref = imp_oracle.Oracle(kind, cfg(9), sd).forward(data)
    net = (DGNNS if kind == 'DGNNS' else AdaGMN)(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net(to_cuda(data))
    mism = sum(int((a.cpu() != b).sum()) for a, b in zip(out['indices0'], ref['indices0']))
    diffs = sorted(float((a.cpu() - b).abs().max()) for a, b in zip(out['mscores0'], ref['mscores0']))
    assert mism == 0
    assert diffs[len(diffs) // 2] < 1e-4, diffs       # median over iterations (a single near-tie flip may spike one)


def test_pair_feeder_stages_batches_in_order():
    """PairFeeder (host -> device double buffering in front of the public API): batches come out intact and in order,
    shape-only entries pass through, and over-staging is refused."""
    from imp_release_b200.feeder import PairFeeder
    feeder = PairFeeder('cuda', depth=2)
    batches = [{'descriptors0': torch.full((2, 50, 256), float(i)).pin_memory(), 'scores0': torch.arange(100.).view(2, 50).pin_memory() + i,
                'image0': torch.zeros(1, 1, 480, 640)} for i in range(5)]
    feeder.stage(batches[0])
    got = []
    for i in range(5):
        d = feeder.next()
        if i + 1 < 5:
            feeder.stage(batches[i + 1])           # lands in the other buffer while batch i is in use
            with pytest.raises(RuntimeError):      # a third batch would overwrite the one in use
                feeder.stage(batches[i + 1])
        assert d['image0'].shape == (1, 1, 480, 640) and not d['image0'].is_cuda
        got.append((d['descriptors0'] * 2).sum().item() / (2 * 50 * 256 * 2) + d['scores0'][0, 0].item())   # uses the slot on the compute stream
    assert got == [2.0 * i for i in range(5)]
    assert PairFeeder.bytes_of(batches[0]) == 2 * 50 * 256 * 4 + 100 * 4
    with pytest.raises(RuntimeError):
        feeder.next()


def test_pose_overlap_async_d2h_of_matches():
    """PoseOverlap with CUDA tensors: the matches of each pair reach the worker through pinned buffers and a CUDA event,
    without synchronising the stream; recycled buffers never leak one pair's matches into another."""
    import numpy as np
    from imp_release_b200.pose_overlap import PoseOverlap
    g = torch.Generator().manual_seed(0)
    pairs = [(torch.randint(-1, 2000, (1000 + 37 * i,), generator=g), torch.rand(1000 + 37 * i, generator=g)) for i in range(20)]
    with PoseOverlap(workers=3, max_pending=4) as po:
        futs = []
        for idx, sc in pairs:
            a, b = idx.cuda(), sc.cuda()
            a2 = a * 1 + 0                                  # something enqueued on the stream before the staging
            futs.append(po.submit(a2, b, lambda i, s: (i.copy(), s.copy())))
        for (idx, sc), f in zip(pairs, futs):
            gi, gs = f.result()
            assert np.array_equal(gi, idx.numpy()) and np.array_equal(gs, sc.numpy())


def test_adagmn_dual_softmax_with_pruning():
    """eval_imp.py --use_dual_softmax with EIMP (with_sinkhorn=False): the dual-softmax scorer must honour the kept
    subsets and feed the pooling with row / column masses (crashed in round 1)."""
    c = cfg(9, with_sinkhorn=False, n_min_tokens=64)
    sd = synth.make_state_dict('AdaGMN', 9, seed=31, bin_score=1.0)
    data = synth.make_pair_batch(seed=32, batch=2, n0=300, n1=280)
    ref = imp_oracle.Oracle('AdaGMN', c, sd).forward(data)
    net = AdaGMN(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net(to_cuda(data))
    cnt, _ = net._kept
    kept = cnt.cpu().tolist()
    assert kept[:2] == ref['kept0'][-1] and kept[2:] == ref['kept1'][-1], (kept, ref['kept0'][-1], ref['kept1'][-1])
    assert min(kept) < 280, 'test vector must actually prune'
    compare(out, ref)


@pytest.mark.parametrize('kind', ['GM', 'AdaGMN'])
def test_run_adapters_gm_adagmn(kind):
    """mode=1 `run` adapters of GM (nets/gm.py:322-364: returns the score matrix 'p') and AdaGMN (nets/adgm.py:607-635:
    matched index pairs)."""
    nl = 3 if kind == 'GM' else 5
    c = cfg(nl, match_threshold=0.2)
    sd = synth.make_state_dict(kind, nl, seed=41, bin_score=1.0)
    data = synth.make_pair_batch(seed=42, batch=1, n0=210, n1=190)
    nk0 = imp_oracle.normalize_keypoints(data['keypoints0'], data['image0'].shape)
    nk1 = imp_oracle.normalize_keypoints(data['keypoints1'], data['image1'].shape)
    rd = {'desc1': data['descriptors0'], 'desc2': data['descriptors1'],
          'x1': torch.cat([nk0, data['scores0'][..., None]], -1), 'x2': torch.cat([nk1, data['scores1'][..., None]], -1)}
    net = (GM if kind == 'GM' else AdaGMN)(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    with torch.no_grad():
        out = net(to_cuda(rd), mode=1)
    orc = imp_oracle.Oracle(kind, c, sd)
    nd = {**data, 'norm_keypoints0': nk0, 'norm_keypoints1': nk1}
    if kind == 'GM':
        ref = orc.produce_matches(nd, p=0.2, only_last=True)
        assert out['p'].shape == ref['scores'][-1].shape
        assert float((out['p'].cpu() - ref['scores'][-1])[:, :-1, :-1].abs().max()) < 1e-3
    else:
        ref = orc.produce_matches(nd, p=0.2)
        ri = ref['indices0'][-1][0]
        idx0 = torch.where(ri >= 0)[0]
        assert torch.equal(out['index0'].cpu(), idx0) and torch.equal(out['index1'].cpu(), ri[idx0])


def test_ragged_pairs_in_one_batch():
    """B200 extension: 'n_keypoints0/1' declare zero-padded inputs, so pairs with different keypoint counts share a
    batch.  Every pair must come out exactly as when it is matched alone."""
    c = cfg(9)
    sd = synth.make_state_dict('DGNNS', 9, seed=7)
    net = DGNNS(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    sizes = [(700, 640), (512, 700), (333, 401), (690, 690)]
    Nmax = 704
    singles = [synth.make_pair_batch(seed=60 + i, batch=1, n0=a, n1=b) for i, (a, b) in enumerate(sizes)]
    batch = {'image0': singles[0]['image0'], 'image1': singles[0]['image1']}
    for k, width in (('descriptors', 256), ('keypoints', 2), ('scores', None)):
        for s in '01':
            shape = (len(sizes), Nmax) + ((width,) if width else ())
            t = torch.zeros(shape)
            for i, d in enumerate(singles):
                t[i, :d[k + s].shape[1]] = d[k + s][0]
            batch[k + s] = t
    batch['n_keypoints0'] = torch.tensor([a for a, _ in sizes], dtype=torch.int32)
    batch['n_keypoints1'] = torch.tensor([b for _, b in sizes], dtype=torch.int32)
    with torch.no_grad():
        out = net.produce_matches(to_cuda(batch), p=0.2, only_last=True)
        i_b, m_b = out['indices0'][-1].cpu(), out['mscores0'][-1].cpu()
        for i, (d, (a, b)) in enumerate(zip(singles, sizes)):
            o = net.produce_matches(to_cuda(d), p=0.2, only_last=True)
            assert torch.equal(i_b[i, :a], o['indices0'][-1][0].cpu()), f'pair {i}'
            assert float((m_b[i, :a] - o['mscores0'][-1][0].cpu()).abs().max()) < 1e-5
            assert bool((i_b[i, a:] == -1).all()) and float(m_b[i, a:].abs().max()) == 0.0
    ref = imp_oracle.Oracle('DGNNS', c, sd).produce_matches(singles[2], p=0.2, only_last=True)
    assert torch.equal(i_b[2, :sizes[2][0]], ref['indices0'][-1][0])


def test_latency_matcher_bucketed_graph_replay():
    """One pair per call with ragged keypoint counts (eval/eval_imp.py:155-173): bucketed static shapes + CUDA-graph replay
    with several pairs in flight must reproduce the eager model (identical matches), with one graph per (slot, 128-bucket)."""
    from imp_release_b200.graphed import LatencyMatcher
    c = cfg(15)
    sd = synth.make_state_dict('DGNNS', 15, seed=7)
    net = DGNNS(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    sizes = [(700, 640), (641, 700), (760, 655), (650, 767), (512, 300), (700, 700), (705, 640), (300, 512)]
    pairs = [to_cuda(synth.make_pair_batch(seed=70 + i, batch=1, n0=a, n1=b)) for i, (a, b) in enumerate(sizes)]
    lm = LatencyMatcher(net, slots=2, p=0.2, only_last=True)
    with torch.no_grad():
        tickets = [lm.submit(d) for d in pairs]                 # all in flight before the first result is read
        outs = [lm.result(t) for t in tickets]
        torch.cuda.synchronize()
        for d, o, (a, b) in zip(pairs, outs, sizes):
            e = net.produce_matches(d, p=0.2, only_last=True)
            assert o['indices0'][-1].shape == (1, a)
            assert torch.equal(o['indices0'][-1], e['indices0'][-1])
            # (the padded problem splits the Sinkhorn rows over CTAs differently: column sums round differently)
            assert float((o['mscores0'][-1] - e['mscores0'][-1]).abs().max()) < 1e-5
    assert lm.captures <= 2 * 3                                  # buckets 512, 768 (+ 640 -> 640): at most 3 per slot
    ref = imp_oracle.Oracle('DGNNS', c, sd).produce_matches({k: v.cpu() for k, v in pairs[2].items()}, only_last=True)
    assert torch.equal(outs[2]['indices0'][-1].cpu(), ref['indices0'][-1])


def test_evaluate_pairs_overlaps_gpu_matching_with_host_pose():
    """the evaluation loop with several pairs in flight on the GPU (LatencyMatcher) and the
    reference's host-side pose step running in worker threads on asynchronously copied matches (PoseOverlap).  Results
    come back in order and equal the serial loop."""
    import numpy as np
    from imp_release_b200 import host_pose
    from imp_release_b200.graphed import LatencyMatcher
    c = cfg(9)
    sd = synth.make_state_dict('DGNNS', 9, seed=7)
    net = DGNNS(c); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    scenes = [synth.make_scene_pair(100 + i, 600 + 20 * i, 640 - 10 * i) for i in range(6)]
    feed = [{k: (v.cuda() if torch.is_tensor(v) and k != 'perm' and not k.startswith('image') else v) for k, v in s.items()} for s in scenes]
    calls = []

    def pose_fn(k0, k1, K0, K1, th):
        calls.append(len(k0))
        return host_pose.estimate_pose(k0, k1, K0, K1, th)
    lm = LatencyMatcher(net, slots=2, p=0.2, only_last=True)
    res = host_pose.evaluate_pairs(lm, feed, pose_fn=pose_fn, workers=2)
    assert len(res) == 6 and len(calls) == 6
    with torch.no_grad():
        for d, r in zip(feed, res):
            out = net.produce_matches(d, p=0.2, only_last=True)
            i0 = out['indices0'][-1][0].cpu().numpy()
            n = int((i0 > -1).sum())
            assert n in calls
            if r is not None:
                assert r[3].shape == (n,) and r[1].shape == (3, 3)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs in one process')
def test_model_on_second_gpu_while_first_is_current():
    """One process driving two GPUs: a model that lives on cuda:1 is called while cuda:0 is the current device.  The entry
    points switch to the model's device (ops.on_model_device) and the launchers configure their kernels per device."""
    from imp_release_b200 import DGNNS
    nl = 3
    cfg = dict(n_layers=nl, GNN_layers=['self', 'cross'] * nl, norm_fn='in', ac_fn='relu', sinkhorn_iterations=20,
               with_sinkhorn=True, descriptor_dim=256)
    sd = synth.make_state_dict('DGNNS', nl, seed=7)
    data = synth.make_pair_batch(seed=4, batch=2, n0=300, n1=260)
    outs = []
    torch.cuda.set_device(0)
    for dev in ('cuda:0', 'cuda:1'):
        net = DGNNS(cfg)
        net.load_state_dict(sd)
        net = net.to(dev).eval()
        with torch.no_grad():
            o = net({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()})
        assert torch.cuda.current_device() == 0
        assert o['indices0'][-1].device == torch.device(dev)
        outs.append(o)
    assert torch.equal(outs[0]['indices0'][-1].cpu(), outs[1]['indices0'][-1].cpu())
    assert float((outs[0]['mscores0'][-1].cpu() - outs[1]['mscores0'][-1].cpu()).abs().max()) < 1e-5
