"""CPU: the oracle restatement against the committed outputs of the UNMODIFIED reference (tests/golden/, generated
by tests/golden/make_golden.py in the build container).  Indices bit-exact, scores <= 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import imp_oracle, synth
from tests.golden.make_golden import CASES, cfg

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_outputs.npz'))


@pytest.mark.parametrize('name', list(CASES))
def test_model_outputs_match_reference(name):
    kind, nl, wseed, bin_score, dseed, B, n0, n1, kw = CASES[name]
    sd = synth.make_state_dict(kind, nl, seed=wseed, bin_score=bin_score)
    data = synth.make_pair_batch(seed=dseed, batch=B, n0=n0, n1=n1)
    assert synth.state_dict_checksum(sd) == pytest.approx(float(G[f'{name}/weights_checksum']), rel=1e-12), 'RNG drift'
    assert sum(synth.tensor_checksum(v) for k, v in sorted(data.items())) == pytest.approx(float(G[f'{name}/data_checksum']), rel=1e-12)
    orc = imp_oracle.Oracle(kind, cfg(nl), sd)
    with torch.no_grad():
        out = orc.produce_matches(data, **kw) if kw else orc.forward(data)
    i0 = torch.stack(out['indices0']).numpy()
    m0 = torch.stack(out['mscores0']).numpy()
    assert np.array_equal(i0, G[f'{name}/indices0'])
    assert np.abs(m0 - G[f'{name}/mscores0']).max() < 1e-5
    if f'{name}/scores_last_shape' in G:
        s = out['scores'][-1]
        assert list(s.shape) == G[f'{name}/scores_last_shape'].tolist()
        ref = G[f'{name}/scores_last_rowsum']
        assert np.abs(s.sum(-1).numpy() - ref).max() / np.abs(ref).max() < 1e-5
        ref = G[f'{name}/scores_last_colsum']
        assert np.abs(s.sum(-2).numpy() - ref).max() / np.abs(ref).max() < 1e-5


@pytest.mark.parametrize('kind,bin_score', [('DGNNS', 1.0), ('AdaGMN', 6.0)])
def test_layer_api_matches_reference(kind, bin_score):
    nl, n0, n1 = 9, 330, 300
    tag = f'layerapi_{kind.lower()}'
    sd = synth.make_state_dict(kind, nl, seed=11, bin_score=bin_score)
    data = synth.make_pair_batch(seed=12, batch=1, n0=n0, n1=n1)
    m = imp_oracle.Oracle(kind, cfg(nl), sd)
    with torch.no_grad():
        nk0 = imp_oracle.normalize_keypoints(data['keypoints0'], data['image0'].shape)
        nk1 = imp_oracle.normalize_keypoints(data['keypoints1'], data['image1'].shape)
        e0, e1 = m.encode_keypoint(nk0, nk1, data['scores0'], data['scores1'])
        d0 = data['descriptors0'].transpose(1, 2) + e0
        d1 = data['descriptors1'].transpose(1, 2) + e1
        for it in range(4):
            d0, d1 = m.forward_one_layer(d0, d1, None, None, 2 * it)
            d0, d1 = m.forward_one_layer(d0, d1, None, None, 2 * it + 1)
        dist = m.compute_distance(d0, d1, layer_id=3)
        score = m.compute_score(dist, m.bin_score, m.sinkhorn_iterations)
        i0, i1, m0, m1 = m.compute_matches(score, p=0.1)
    assert np.abs(d0.numpy()[:, :, ::7] - G[f'{tag}/desc0_it3']).max() < 2e-4
    assert np.array_equal(i0.numpy(), G[f'{tag}/indices0']) and np.array_equal(i1.numpy(), G[f'{tag}/indices1'])
    assert np.abs(m0.numpy() - G[f'{tag}/mscores0']).max() < 1e-5
    if kind == 'AdaGMN':
        ids0, ids1 = m.pool(pred_score=score, prob00=m.self_prob0, prob01=m.cross_prob0, prob11=m.self_prob1,
                            prob10=m.cross_prob1, mscore_th=0.2, uncertainty_ratio=1.0)
        assert np.array_equal(ids0.numpy(), G[f'{tag}/pool_ids0']) and np.array_equal(ids1.numpy(), G[f'{tag}/pool_ids1'])
        assert len(ids0) < n0 and len(ids1) < n1


def test_free_functions_known_answers():
    M = torch.from_numpy(G['fn/sink_in'])
    bs = torch.tensor(0.8)
    for it, key in ((20, 'fn/sink_out20'), (0, 'fn/sink_out0')):
        out = imp_oracle.sink_algorithm(M, bs, it).numpy()
        assert np.abs(out - G[key]).max() / np.abs(G[key]).max() < 1e-6
    assert np.abs(imp_oracle.dual_softmax(M, bs).numpy() - G['fn/dual_softmax']).max() < 1e-6
    i0, i1, m0, m1 = imp_oracle.compute_matches(torch.from_numpy(G['fn/ties_in']), 0.2)
    assert np.array_equal(i0.numpy(), G['fn/ties_i0']) and np.array_equal(i1.numpy(), G['fn/ties_i1'])
    assert np.array_equal(m0.numpy(), G['fn/ties_m0']) and np.array_equal(m1.numpy(), G['fn/ties_m1'])
    kp = torch.from_numpy(G['fn/normkp_in'])
    assert np.abs(imp_oracle.normalize_keypoints(kp, (1, 1, 480, 640)).numpy() - G['fn/normkp_out']).max() < 1e-6


def test_sinkhorn_against_independent_fp64_log_domain():
    """Cross-check of the probability-domain recurrence with an independent fp64 log-domain derivation
    (the form of nets/superglue.py:180-209 with this model's marginals): identical assignment, tiny score gap."""
    g = torch.Generator().manual_seed(3)
    M = torch.randn(1, 60, 70, generator=g) * 3
    bs = torch.tensor(1.0)
    out = imp_oracle.sink_algorithm(M, bs, 20)
    Z = torch.log_softmax(imp_oracle.pad_dustbin(M, bs).double(), -1)
    logr = torch.zeros(1, 61, dtype=torch.float64); logr[:, -1] = np.log(61)
    logc = torch.zeros(1, 71, dtype=torch.float64); logc[:, -1] = np.log(71)
    f = torch.zeros_like(logr); gg = torch.zeros_like(logc)
    for _ in range(20):
        f = logr - torch.logsumexp(Z + gg[:, None, :], -1)
        gg = logc - torch.logsumexp(Z + f[:, :, None], -2)
    ref = torch.exp(Z + f[:, :, None] + gg[:, None, :])
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 1e-4
    assert torch.equal(out[:, :-1, :-1].argmax(-1), ref[:, :-1, :-1].argmax(-1))


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json headline shape (N = 2000, 9 iterations) and the iterative drivers of eval/matching.py
from tests.golden.make_golden import MATCHING_CASES, MATCHING_NI, N2000_CASES, run_matching_case  # noqa: E402
from tests import matching_driver as md  # noqa: E402
from oracle import refimport  # noqa: E402

G2K = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_n2000.npz'))
GM_ = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_matching.npz'))


@pytest.mark.parametrize('name', list(N2000_CASES))
def test_headline_shape_oracle_matches_reference(name):
    kind, nl, wseed, bin_score, dseed, B, n0, n1 = N2000_CASES[name]
    sd = synth.make_state_dict(kind, nl, seed=wseed, bin_score=bin_score)
    data = synth.make_pair_batch(seed=dseed, batch=B, n0=n0, n1=n1)
    assert synth.state_dict_checksum(sd) == pytest.approx(float(G2K[f'{name}/weights_checksum']), rel=1e-12), 'RNG drift'
    assert sum(synth.tensor_checksum(v) for k, v in sorted(data.items())) == pytest.approx(float(G2K[f'{name}/data_checksum']), rel=1e-12)
    with torch.no_grad():
        out = imp_oracle.Oracle(kind, cfg(nl), sd).forward(data)
    assert np.array_equal(torch.stack(out['indices0']).numpy(), G2K[f'{name}/indices0'].astype(np.int64))
    assert np.abs(torch.stack(out['mscores0']).numpy() - G2K[f'{name}/mscores0']).max() < 1e-5


def check_trace(name, tr, ret, driver, pose, G=GM_, tol=1e-5, exact_pool=True):
    """A driver run against the trace the unmodified reference driver + reference model produced."""
    assert len(tr.events) == int(G[f'{name}/n_events']), 'the driver took a different path (number of scorings / pools)'
    for k, e in enumerate(tr.events):
        assert e[0] == str(G[f'{name}/ev{k}/kind']) and e[1] == pytest.approx(float(G[f'{name}/ev{k}/p']), abs=1e-7), f'event {k}'
        if e[0] == 'matches':
            assert np.array_equal(e[2], G[f'{name}/ev{k}/a']), f'event {k}: {(e[2] != G[f"{name}/ev{k}/a"]).sum()} match indices differ'
            assert np.abs(e[3] - G[f'{name}/ev{k}/b']).max() < tol, f'event {k}: match scores'
        else:
            assert np.array_equal(e[2], G[f'{name}/ev{k}/a']) and np.array_equal(e[3], G[f'{name}/ev{k}/b']), f'event {k}: kept ids differ'
    i0, m0, n_it = (ret[0], ret[1], ret[4]) if driver == 'matching_iterative' else (ret[4], ret[5], ret[8])
    assert int(n_it) == int(G[f'{name}/ret_iterations'])
    assert np.array_equal(np.asarray(i0), G[f'{name}/ret_indices0'])
    assert np.abs(np.asarray(m0) - G[f'{name}/ret_mscores0']).max() < tol
    assert pose.n_matches == G[f'{name}/pose_calls'].tolist()
    if driver == 'matching_iterative_uncertainty':
        assert [len(ret[0]), len(ret[1])] == G[f'{name}/ret_n_pts'].tolist()


@pytest.mark.parametrize('name', list(MATCHING_CASES))
def test_restated_drivers_on_oracle_match_reference_drivers(name):
    """tests/matching_driver.py (restated eval/matching.py loops) driving the ORACLE reproduces what the unmodified
    drivers saw and returned on the reference model: pins the restated drivers and the oracle's per-layer API
    (incl. pool and the caller-side compaction) on any box."""
    case = MATCHING_CASES[name]
    driver, kind, wseed, bin_score = case[:4]
    sd = synth.make_state_dict(kind, MATCHING_NI, seed=wseed, bin_score=bin_score)
    m = imp_oracle.Oracle(kind, cfg(MATCHING_NI), sd)
    pose = md.PoseStub()
    tr, ret = run_matching_case(case, m, getattr(md, driver), pose, estimate_pose=pose,
                                normalize_keypoints=imp_oracle.normalize_keypoints)
    check_trace(name, tr, ret, driver, pose, tol=5e-5)       # fp32 summation-order drift over up to 30 layers


@pytest.mark.skipif(not refimport.available(), reason='needs /root/reference (build container only)')
def test_restated_drivers_equal_unmodified_drivers_on_reference_model():
    """Same reference model under the restated loop and under the fixture made by the UNMODIFIED eval/matching.py."""
    ns = refimport.load()
    try:
        for name in ('mi_dgnns_stop', 'miu_adagmn_stop'):
            case = MATCHING_CASES[name]
            driver, kind, wseed, bin_score = case[:4]
            sd = synth.make_state_dict(kind, MATCHING_NI, seed=wseed, bin_score=bin_score)
            m = {'DGNNS': ns.DGNNS, 'AdaGMN': ns.AdaGMN}[kind](cfg(MATCHING_NI)).eval()
            m.load_state_dict(sd, strict=True)
            pose = md.PoseStub()
            tr, ret = run_matching_case(case, m, getattr(md, driver), pose, estimate_pose=pose,
                                        normalize_keypoints=ns.layers.normalize_keypoints)
            check_trace(name, tr, ret, driver, pose, tol=1e-7)
    finally:
        import sys
        for k in [k for k in sys.modules if k == 'nets' or k.startswith('nets.') or k == 'eval' or k.startswith('eval.') or k == 'tools' or k.startswith('tools.')]:
            del sys.modules[k]


# ---------------------------------------------------------------------------------------------------------------
# SuperPoint front-end: oracle/superpoint_oracle.py against the unmodified reference class
from oracle import superpoint_oracle as spo  # noqa: E402
from tests.golden.make_golden_superpoint import CASES as SP_CASES, DEFAULT as SP_DEFAULT, probe_dirs  # noqa: E402

SPG = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_superpoint.npz'))


@pytest.mark.parametrize('name', list(SP_CASES))
def test_superpoint_oracle_matches_reference(name):
    wseed, iseed, H, W, B, over = SP_CASES[name]
    sd = spo.make_state_dict(wseed)
    img = spo.make_image(iseed, H, W, B)
    assert sum(float(v.double().abs().sum()) for v in sd.values()) == pytest.approx(float(SPG[f'{name}/weights_checksum']), rel=1e-12)
    assert float(img.double().sum()) == pytest.approx(float(SPG[f'{name}/image_checksum']), rel=1e-12), 'RNG drift'
    with torch.no_grad():
        scores, _ = spo.dense(sd, img)
        out = spo.forward(sd, img, {**SP_DEFAULT, **over})
    assert np.abs(scores[:, ::8, ::8].numpy() - SPG[f'{name}/dense_scores_8x']).max() < 1e-7
    for b in range(B):
        assert np.array_equal(out['keypoints'][b].numpy().astype(np.int32), SPG[f'{name}/{b}/keypoints'])
        assert np.abs(out['scores'][b].numpy() - SPG[f'{name}/{b}/scores']).max() < 1e-7
        d = out['descriptors'][b]
        assert np.abs(d[:, :48].numpy() - SPG[f'{name}/{b}/descriptors_head']).max() < 1e-6
        assert np.abs((d.t() @ probe_dirs()).numpy() - SPG[f'{name}/{b}/descriptor_probes']).max() < 1e-5
