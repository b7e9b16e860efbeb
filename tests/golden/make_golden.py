"""Generate the committed golden fixtures by running the UNMODIFIED reference (feixue94/imp-release) on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Inputs and weights are regenerated from seeds by oracle/synth.py (checksums are stored so RNG drift is detected);
outputs are what the reference classes nets.gm.GM / nets.gms.DGNNS / nets.adgm.AdaGMN return (shims: see
oracle/refimport.py).  Everything is small enough to commit (< 1 MB).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refimport, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cfg(nl, **kw):
    c = dict(n_layers=nl, GNN_layers=['self', 'cross'] * nl, norm_fn='in', ac_fn='relu', sinkhorn_iterations=20,
             with_sinkhorn=True, descriptor_dim=256, n_min_tokens=256)
    c.update(kw)
    return c


CASES = {
    # name: (kind, n_layers, weight seed, bin_score, data seed, batch, n0, n1, call kwargs)
    'gm_config1': ('GM', 3, 5, 1.0, 0, 1, 512, 512, {}),                       # BASELINE.json configs[0]
    'dgnns_all': ('DGNNS', 9, 7, 1.0, 3, 2, 300, 260, {'only_last': False}),
    'dgnns_last': ('DGNNS', 9, 7, 1.0, 4, 1, 420, 500, {'only_last': True}),
    'adagmn_prune': ('AdaGMN', 9, 7, 6.0, 8, 2, 400, 380, {}),
}


def main():
    assert refimport.available(), 'reference tree missing'
    ns = refimport.load()
    cls = {'GM': ns.GM, 'DGNNS': ns.DGNNS, 'AdaGMN': ns.AdaGMN}
    torch.manual_seed(0)
    blob = {}
    for name, (kind, nl, wseed, bin_score, dseed, B, n0, n1, kw) in CASES.items():
        sd = synth.make_state_dict(kind, nl, seed=wseed, bin_score=bin_score)
        data = synth.make_pair_batch(seed=dseed, batch=B, n0=n0, n1=n1)
        m = cls[kind](cfg(nl)).eval()
        m.load_state_dict(sd, strict=True)          # proves the state_dict spec of oracle/synth.py
        with torch.no_grad():
            out = m.produce_matches(data, **kw) if kw else m(data)
        blob[f'{name}/indices0'] = torch.stack(out['indices0']).numpy()
        blob[f'{name}/mscores0'] = torch.stack(out['mscores0']).numpy()
        blob[f'{name}/weights_checksum'] = np.float64(synth.state_dict_checksum(sd))
        blob[f'{name}/data_checksum'] = np.float64(sum(synth.tensor_checksum(v) for k, v in sorted(data.items())))
        if 'scores' in out:
            s = out['scores'][-1]
            blob[f'{name}/scores_last_shape'] = np.array(s.shape)
            blob[f'{name}/scores_last_rowsum'] = s.sum(-1).numpy()
            blob[f'{name}/scores_last_colsum'] = s.sum(-2).numpy()
            blob[f'{name}/scores_last_diag'] = s[:, torch.arange(min(s.shape[1:])), torch.arange(min(s.shape[1:]))].numpy()
        print(name, 'matches per iteration', [(i >= 0).sum().item() for i in out['indices0']])

    # per-layer API of DGNNS / AdaGMN (eval/matching.py call sequence) incl. pool
    for kind, bin_score in (('DGNNS', 1.0), ('AdaGMN', 6.0)):
        nl, n0, n1 = 9, 330, 300
        sd = synth.make_state_dict(kind, nl, seed=11, bin_score=bin_score)
        data = synth.make_pair_batch(seed=12, batch=1, n0=n0, n1=n1)
        m = cls[kind](cfg(nl)).eval()
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            nk0 = ns.layers.normalize_keypoints(data['keypoints0'], data['image0'].shape)
            nk1 = ns.layers.normalize_keypoints(data['keypoints1'], data['image1'].shape)
            e0, e1 = m.encode_keypoint(nk0, nk1, data['scores0'], data['scores1'])
            d0 = data['descriptors0'].transpose(1, 2) + e0
            d1 = data['descriptors1'].transpose(1, 2) + e1
            for it in range(4):
                d0, d1 = m.forward_one_layer(d0, d1, None, None, 2 * it)
                d0, d1 = m.forward_one_layer(d0, d1, None, None, 2 * it + 1)
            dist = m.compute_distance(d0, d1, layer_id=3)
            score = m.compute_score(dist, m.bin_score, m.sinkhorn_iterations)
            i0, i1, m0, m1 = m.compute_matches(score, p=0.1)
            tag = f'layerapi_{kind.lower()}'
            blob[f'{tag}/desc0_it3'] = d0.numpy().astype(np.float32)[:, :, ::7]
            blob[f'{tag}/indices0'] = i0.numpy(); blob[f'{tag}/indices1'] = i1.numpy()
            blob[f'{tag}/mscores0'] = m0.numpy(); blob[f'{tag}/mscores1'] = m1.numpy()
            blob[f'{tag}/weights_checksum'] = np.float64(synth.state_dict_checksum(sd))
            if kind == 'AdaGMN':
                ids0, ids1 = m.pool(pred_score=score, prob00=m.self_prob0, prob01=m.cross_prob0, prob11=m.self_prob1,
                                    prob10=m.cross_prob1, mscore_th=0.2, uncertainty_ratio=1.0)
                blob[f'{tag}/pool_ids0'] = ids0.numpy() if ids0 is not None else np.array([-1])
                blob[f'{tag}/pool_ids1'] = ids1.numpy() if ids1 is not None else np.array([-1])
                print(tag, 'pool kept', None if ids0 is None else len(ids0), None if ids1 is None else len(ids1))

    # known answers for the free functions
    g = torch.Generator().manual_seed(99)
    M = torch.randn(2, 37, 29, generator=g) * 2
    bs = torch.tensor(0.8)
    r = torch.ones(2, 38); r[:, -1] = 38
    c = torch.ones(2, 30); c[:, -1] = 30
    Ma = torch.cat([torch.cat([M, bs.expand(2, 37, 1)], -1), bs.expand(2, 1, 30)], -2)
    blob['fn/sink_in'] = M.numpy()
    blob['fn/sink_out20'] = ns.layers.sinkhorn(Ma, r, c, 20).numpy()
    blob['fn/sink_out0'] = ns.layers.sinkhorn(Ma, r, c, 0).numpy()
    blob['fn/dual_softmax'] = ns.layers.dual_softmax(M, bs).numpy()
    ties = torch.zeros(1, 6, 5)
    ties[0, 1, 2] = ties[0, 3, 2] = 0.9      # column tie -> lowest row wins
    ties[0, 4, 0] = ties[0, 4, 3] = 0.7      # row tie -> lowest column wins
    gm = cls['GM'](cfg(1)).eval()
    ti0, ti1, tm0, tm1 = gm.compute_matches(ties, p=0.2)
    blob['fn/ties_in'] = ties.numpy()
    blob['fn/ties_i0'] = ti0.numpy(); blob['fn/ties_i1'] = ti1.numpy()
    blob['fn/ties_m0'] = tm0.numpy(); blob['fn/ties_m1'] = tm1.numpy()
    kp = torch.rand(1, 9, 2, generator=g) * torch.tensor([640., 480.])
    blob['fn/normkp_in'] = kp.numpy()
    blob['fn/normkp_out'] = ns.layers.normalize_keypoints(kp, (1, 1, 480, 640)).numpy()

    path = os.path.join(OUT, 'reference_outputs.npz')
    np.savez_compressed(path, **blob)
    print('wrote', path, os.path.getsize(path), 'bytes')


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json headline shape (N = 2000, 9 iterations): the unmodified reference on two pairs; the GPU test runs
# these two pairs inside a 64-pair batch (tests/test_golden_gpu.py::test_headline_shape_matches_reference).
N2000_CASES = {
    # name: (kind, n_layers, weight seed, bin_score, data seed, batch, n0, n1)
    'dgnns_n2000': ('DGNNS', 9, 7, 1.0, 41, 2, 2000, 2000),          # configs[1]: model(data), Sinkhorn every iteration
    'adagmn_n2000': ('AdaGMN', 9, 7, 8.0, 42, 2, 2000, 2000),        # configs[2]: pruning 2000 -> ~600
}


def main_n2000():
    ns = refimport.load()
    cls = {'GM': ns.GM, 'DGNNS': ns.DGNNS, 'AdaGMN': ns.AdaGMN}
    blob = {}
    for name, (kind, nl, wseed, bin_score, dseed, B, n0, n1) in N2000_CASES.items():
        sd = synth.make_state_dict(kind, nl, seed=wseed, bin_score=bin_score)
        data = synth.make_pair_batch(seed=dseed, batch=B, n0=n0, n1=n1)
        m = cls[kind](cfg(nl)).eval()
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out = m(data)
        i0 = torch.stack(out['indices0']).numpy()
        assert i0.max() < 32768
        blob[f'{name}/indices0'] = i0.astype(np.int16)
        blob[f'{name}/mscores0'] = torch.stack(out['mscores0']).numpy()
        blob[f'{name}/weights_checksum'] = np.float64(synth.state_dict_checksum(sd))
        blob[f'{name}/data_checksum'] = np.float64(sum(synth.tensor_checksum(v) for k, v in sorted(data.items())))
        if 'scores' in out and out['scores'][-1] is not None:
            blob[f'{name}/scores_last_shape'] = np.array(out['scores'][-1].shape)
        print(name, 'matches per iteration', [(i >= 0).sum().item() for i in out['indices0']])
    path = os.path.join(OUT, 'reference_n2000.npz')
    np.savez_compressed(path, **blob)
    print('wrote', path, os.path.getsize(path), 'bytes')


# ---------------------------------------------------------------------------------------------------------------
# The reference's own iterative drivers (eval/matching.py:16-276, UNMODIFIED) on the reference model, with the pose
# solver replaced by the deterministic tests/matching_driver.py::PoseStub.  Recorded: every compute_matches / pool
# result the driver saw, plus what it returned.
MATCHING_CASES = {
    # name: (driver, kind, weight seed, bin_score, data seed, n0, n1, stop criteria, with_uncertainty)
    'mi_dgnns_stop': ('matching_iterative', 'DGNNS', 23, 1.0, 51, 800, 760, {'match': 0.7, 'pose': 1.5}, False),
    'mi_dgnns_full': ('matching_iterative', 'DGNNS', 23, 1.0, 52, 640, 700, {}, False),
    'miu_adagmn_stop': ('matching_iterative_uncertainty', 'AdaGMN', 29, 6.0, 53, 800, 760, {'match': 0.7, 'pose': 1.5}, True),
    'miu_adagmn_full': ('matching_iterative_uncertainty', 'AdaGMN', 29, 6.0, 54, 700, 780, {}, False),
}
MATCHING_NI = 15


def run_matching_case(case, model, driver_fn, pose, **extra):
    from tests import matching_driver as md
    driver, kind, wseed, bin_score, dseed, n0, n1, stop, unc = case
    data = md.make_driver_data(synth, dseed, n0, n1)
    tr = md.Trace(model)
    kw = dict(data=data, model=model, nI=MATCHING_NI, match_ratio=0.1, min_kpts=25, error_th=1.0, stop_criteria=stop, **extra)
    if driver == 'matching_iterative_uncertainty':
        kw['with_uncertainty'] = unc
    with torch.no_grad():
        ret = driver_fn(**kw)
    tr.close()
    return tr, ret


def main_matching():
    from tests import matching_driver as md
    ns = refimport.load()
    ref_matching = refimport.load_matching(ns)
    cls = {'DGNNS': ns.DGNNS, 'AdaGMN': ns.AdaGMN}
    blob = {}
    for name, case in MATCHING_CASES.items():
        driver, kind, wseed, bin_score = case[:4]
        sd = synth.make_state_dict(kind, MATCHING_NI, seed=wseed, bin_score=bin_score)
        m = cls[kind](cfg(MATCHING_NI)).eval()
        m.load_state_dict(sd, strict=True)
        pose = md.PoseStub()
        ref_matching.estimate_pose = pose                      # the only patch: the (randomised) RANSAC solver
        tr, ret = run_matching_case(case, m, getattr(ref_matching, driver), pose)
        blob.update(tr.to_blob(name))
        i0, m0, n_it = (ret[0], ret[1], ret[4]) if driver == 'matching_iterative' else (ret[4], ret[5], ret[8])
        blob[f'{name}/ret_indices0'] = np.asarray(i0).astype(np.int32)
        blob[f'{name}/ret_mscores0'] = np.asarray(m0).astype(np.float32)
        blob[f'{name}/ret_iterations'] = np.int64(n_it)
        blob[f'{name}/pose_calls'] = np.array(pose.n_matches, dtype=np.int64)
        if driver == 'matching_iterative_uncertainty':
            blob[f'{name}/ret_n_pts'] = np.array([len(ret[0]), len(ret[1])])
        print(name, 'events', len(tr.events), 'pose calls', pose.n_matches, 'stopped after', n_it,
              'pool sizes', [(len(e[2]), len(e[3])) for e in tr.events if e[0] == 'pool'])
    path = os.path.join(OUT, 'reference_matching.npz')
    np.savez_compressed(path, **blob)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    assert refimport.available(), 'reference tree missing'
    which = sys.argv[1:] or ['base', 'n2000', 'matching']
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if 'base' in which:
        main()
    if 'n2000' in which:
        main_n2000()
    if 'matching' in which:
        main_matching()
