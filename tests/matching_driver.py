"""TEST INFRASTRUCTURE: a restatement of the reference's iterative drivers, eval/matching.py:16-123
(``matching_iterative``) and :126-276 (``matching_iterative_uncertainty``), so that the GPU box (which has no
/root/reference) can replay the exact call sequence those drivers issue against a model -- the drop-in classes of
``dropin/nets`` there, the reference's own classes in the build container.

``tests/test_oracle_golden.py::test_restated_drivers_equal_reference_drivers`` runs the UNMODIFIED drivers and these
restatements side by side on the reference model (CPU, build container) and requires identical traces; the GPU tests
then run the restatements over the CUDA classes and compare with the traces the unmodified drivers produced
(tests/golden/make_golden.py -> tests/golden/reference_matching.npz).

The pose solver is replaced in BOTH by ``PoseStub`` (deterministic, defined here): RANSAC is randomised and is host
code above the boundary (eval/pose_estimation.py:92-115).
"""
from __future__ import annotations

import numpy as np
import torch

VALID_ITS = [3, 5, 7, 9, 11, 13, 14]          # eval/matching.py:45, :155


class PoseStub:
    """Deterministic stand-in for eval.pose_estimation.estimate_pose (same signature, same return convention).
    The k-th call returns a rotation of 20/(k+1) degrees about z and the matching unit translation, so consecutive
    estimates converge and the 'pose' stop criterion (eval/matching.py:113-121) fires on a known call; the inlier mask
    is a function of the matched coordinates, so a wrong or re-ordered match changes the output."""

    def __init__(self):
        self.calls = 0
        self.n_matches = []

    def __call__(self, kpts0, kpts1, K0, K1, norm_thresh, conf=0.99999, method=None, mask=None):
        if len(kpts0) < 5:                                   # eval/pose_estimation.py:93-94
            return None
        a = np.deg2rad(20.0 / (self.calls + 1))
        self.calls += 1
        self.n_matches.append(len(kpts0))
        R = np.array([[np.cos(a), -np.sin(a), 0.], [np.sin(a), np.cos(a), 0.], [0., 0., 1.]])
        t = np.array([np.cos(a), np.sin(a), 0.])
        key = np.floor(np.asarray(kpts0)[:, 0]).astype(np.int64) + np.floor(np.asarray(kpts1)[:, 1]).astype(np.int64)
        inliers = (key % 4) != 3
        return np.eye(3), R, t, inliers


def angle_error_mat(R1, R2):                                  # tools/utils.py:425-428
    cos = np.clip((np.trace(np.dot(R1.T, R2)) - 1) / 2, -1., 1.)
    return np.rad2deg(np.abs(np.arccos(cos)))


def angle_error_vec(v1, v2):                                  # tools/utils.py:431-433
    n = np.linalg.norm(v1) * np.linalg.norm(v2)
    return np.rad2deg(np.arccos(np.clip(np.dot(v1, v2) / n, -1.0, 1.0)))


def _matches_np(indices0, mscores0):
    """eval/matching.py:70-83."""
    i0 = indices0[0].cpu().numpy()
    m0 = mscores0[0].cpu().numpy()
    ids0 = np.nonzero(i0 > -1)[0]
    return i0, m0, np.stack([ids0, i0[ids0]], 1) if len(ids0) else np.zeros((0, 2), dtype=np.int64)


def _pose_step(estimate_pose, pts0_cpu, pts1_cpu, pred_matches, K0, K1, error_th, method, it, last_R, last_t):
    """eval/matching.py:88-111: pose from the current matches and its distance to the previous estimate."""
    ret = estimate_pose(kpts0=pts0_cpu[pred_matches[:, 0]], kpts1=pts1_cpu[pred_matches[:, 1]], K0=K0, K1=K1,
                        norm_thresh=error_th, method=method)
    if ret is not None:
        _, R, t, inl = ret
        ratio = np.sum(inl) / pred_matches.shape[0]
    else:
        R, t = None, None
        inl = np.zeros(pred_matches.shape[0], dtype=bool)
        ratio = 0
    if it >= 1:
        dR = angle_error_mat(last_R, R) if last_R is not None and R is not None else np.inf
        dt = angle_error_vec(last_t, t) if last_t is not None and t is not None else np.inf
    else:
        dR, dt = np.inf, np.inf
    return R, t, inl, ratio, np.max([dR, dt])


def matching_iterative(data, model, nI, match_ratio, min_kpts, error_th, stop_criteria, estimate_pose, method=None,
                       normalize_keypoints=None):
    """eval/matching.py:16-123.  Returns (indices0, mscores0, R, t, n_iterations)."""
    pts0, pts1 = data['keypoints0'], data['keypoints1']
    # Q1: the reference tests the misspelt key 'norm_keypoint0', so it always normalises here
    nk0 = normalize_keypoints(kpts=pts0, image_shape=data['image0'].shape)
    nk1 = normalize_keypoints(kpts=pts1, image_shape=data['image1'].shape)
    desc0, desc1 = data['descriptors0'].transpose(1, 2), data['descriptors1'].transpose(1, 2)
    last_R = last_t = None
    pred_score = None
    for it in range(nI):
        if it == 0:
            enc0, enc1 = model.encode_keypoint(norm_kpts0=nk0, norm_kpts1=nk1, scores0=data['scores0'],
                                               scores1=data['scores1'])
            desc0, desc1 = desc0 + enc0, desc1 + enc1
        desc0, desc1 = model.forward_one_layer(desc0=desc0, desc1=desc1, M0=None, M1=None, layer_i=it * 2)
        desc0, desc1 = model.forward_one_layer(desc0=desc0, desc1=desc1, M0=None, M1=None, layer_i=it * 2 + 1)
        if it not in VALID_ITS:
            continue
        pred_dist = model.compute_distance(desc0=desc0, desc1=desc1, layer_id=it)
        pred_score = model.compute_score(dist=pred_dist, dustbin=model.bin_score, iteration=model.sinkhorn_iterations)
        indices0, indices1, mscores0, mscores1 = model.compute_matches(scores=pred_score, p=match_ratio)
        if torch.sum(indices0 > -1) < min_kpts:
            last_R = last_t = None
            continue
        i0, m0, pred_matches = _matches_np(indices0, mscores0)
        _ = pred_score.cpu().numpy()[0, :-1, :-1]             # the reference materialises it (eval/matching.py:72)
        if pred_matches.shape[0] == 0:
            continue
        R, t, inl, _, pose_diff = _pose_step(estimate_pose, data['pts0_cpu'], data['pts1_cpu'], pred_matches, data['K0'],
                                             data['K1'], error_th, method, it, last_R, last_t)
        last_R, last_t = R, t
        if 'pose' in stop_criteria and pose_diff <= stop_criteria['pose']:
            out = np.zeros_like(i0) - 1
            out[pred_matches[inl, 0]] = pred_matches[inl, 1]
            return out, m0, R, t, it + 1
    indices0, _, mscores0, _ = model.compute_matches(scores=pred_score, p=0.2)
    return indices0[0].cpu().numpy(), mscores0[0].cpu().numpy(), None, None, nI


def matching_iterative_uncertainty(data, model, nI, match_ratio, min_kpts, error_th, stop_criteria, estimate_pose,
                                   method=None, with_uncertainty=False, normalize_keypoints=None):
    """eval/matching.py:126-276.  Returns (pts0, pts1, norm_kpts0, norm_kpts1, indices0, mscores0, R, t, n_iterations)."""
    pts0, pts1 = data['keypoints0'], data['keypoints1']
    nk0 = normalize_keypoints(kpts=pts0, image_shape=data['image0'].shape)
    nk1 = normalize_keypoints(kpts=pts1, image_shape=data['image1'].shape)
    desc0, desc1 = data['descriptors0'].transpose(1, 2), data['descriptors1'].transpose(1, 2)
    pts0_cpu, pts1_cpu = data['pts0_cpu'], data['pts1_cpu']
    last_R = last_t = None
    sel0 = sel1 = None
    enc0, enc1 = model.encode_keypoint(norm_kpts0=nk0, norm_kpts1=nk1, scores0=data['scores0'], scores1=data['scores1'])
    desc0, desc1 = desc0 + enc0, desc1 + enc1
    upd0 = upd1 = False
    pred_score = None
    for it in range(nI):
        if upd0:                                              # eval/matching.py:166-169: the CALLER compacts
            desc0 = desc0[:, :, sel0]
            pts0_cpu = pts0_cpu[sel0.cpu().numpy()]
            nk0 = nk0[:, sel0, :]
        if upd1:
            desc1 = desc1[:, :, sel1]
            pts1_cpu = pts1_cpu[sel1.cpu().numpy()]
            nk1 = nk1[:, sel1.cpu(), :] if not nk1.is_cuda else nk1[:, sel1, :]
        desc0, desc1 = model.forward_one_layer(desc0=desc0, desc1=desc1, M0=None, M1=None, layer_i=it * 2)
        desc0, desc1 = model.forward_one_layer(desc0=desc0, desc1=desc1, M0=None, M1=None, layer_i=it * 2 + 1)
        if it not in VALID_ITS:
            upd0 = upd1 = False
            continue
        prob00, prob11, prob01, prob10 = model.self_prob0, model.self_prob1, model.cross_prob0, model.cross_prob1
        pred_dist = model.compute_distance(desc0=desc0, desc1=desc1, layer_id=it)
        pred_score = model.compute_score(dist=pred_dist, dustbin=model.bin_score, iteration=model.sinkhorn_iterations)
        indices0, indices1, mscores0, mscores1 = model.compute_matches(scores=pred_score, p=match_ratio)
        if torch.sum(indices0 > -1) < min_kpts:
            last_R = last_t = None
            continue                                          # (update flags keep their value, like the reference)
        i0, m0, pred_matches = _matches_np(indices0, mscores0)
        _ = pred_score.cpu().numpy()[0, :-1, :-1]
        if pred_matches.shape[0] == 0:
            continue
        R, t, inl, ratio, pose_diff = _pose_step(estimate_pose, pts0_cpu, pts1_cpu, pred_matches, data['K0'], data['K1'],
                                                 error_th, method, it, last_R, last_t)
        last_R, last_t = R, t
        mscore_th = (0.2 if ratio == 0 else 0.2 * ratio) if with_uncertainty else 0.2
        sel0, sel1 = model.pool(pred_score=pred_score, prob00=prob00, prob01=prob01, prob11=prob11, prob10=prob10,
                                mscore_th=mscore_th, uncertainty_ratio=1.0)
        upd0, upd1 = sel0 is not None, sel1 is not None
        if 'pose' in stop_criteria and pose_diff <= stop_criteria['pose']:
            out = np.zeros_like(i0) - 1
            out[pred_matches[inl, 0]] = pred_matches[inl, 1]
            return pts0_cpu, pts1_cpu, nk0[0].cpu().numpy(), nk1[0].cpu().numpy(), out, m0, R, t, it + 1
    indices0, _, mscores0, _ = model.compute_matches(scores=pred_score, p=0.2)
    return (pts0_cpu, pts1_cpu, nk0[0].cpu().numpy(), nk1[0].cpu().numpy(), indices0[0].cpu().numpy(),
            mscores0[0].cpu().numpy(), None, None, nI)


class Trace:
    """Records what a driver made the model produce: every compute_matches result and every pool decision."""

    def __init__(self, model):
        self.events = []
        self._model = model
        cm, pool = model.compute_matches, model.pool

        def compute_matches(scores, p=0.2):
            out = cm(scores=scores, p=p)
            self.events.append(('matches', float(p), out[0][0].cpu().numpy().astype(np.int32), out[2][0].cpu().numpy()))
            return out

        def pool_(**kw):
            s0, s1 = pool(**kw)
            self.events.append(('pool', float(kw.get('mscore_th', 0.1)),
                                np.array([-1]) if s0 is None else s0.cpu().numpy().astype(np.int32),
                                np.array([-1]) if s1 is None else s1.cpu().numpy().astype(np.int32)))
            return s0, s1

        model.compute_matches = compute_matches
        model.pool = pool_

    def close(self):
        del self._model.compute_matches, self._model.pool      # drop the instance attributes -> class methods again

    def to_blob(self, tag):
        blob = {f'{tag}/n_events': np.int64(len(self.events))}
        for k, e in enumerate(self.events):
            blob[f'{tag}/ev{k}/kind'] = np.array(e[0])
            blob[f'{tag}/ev{k}/p'] = np.float64(e[1])
            blob[f'{tag}/ev{k}/a'] = e[2]
            blob[f'{tag}/ev{k}/b'] = e[3]
        return blob


def make_driver_data(synth, seed, n0, n1, device='cpu'):
    """eval/eval_imp.py:59-78 feed_data for the iterative drivers; image0 is [1,H,W,3] there (quirk Q3: the model then
    normalises with height = W, width = 3)."""
    d = synth.make_pair_batch(seed=seed, batch=1, n0=n0, n1=n1)
    data = {k: (v.to(device) if k.startswith(('desc', 'keyp', 'scor')) else v) for k, v in d.items()}
    data['image0'] = torch.zeros(1, 480, 640, 3)
    data['image1'] = torch.zeros(1, 480, 640, 3)
    data['pts0_cpu'] = d['keypoints0'][0].numpy()
    data['pts1_cpu'] = d['keypoints1'][0].numpy()
    data['K0'] = data['K1'] = np.array([[500., 0., 320.], [0., 500., 240.], [0., 0., 1.]])
    data['T_0to1'] = np.hstack([np.eye(3), np.array([[1.], [0.], [0.]])])
    return data
