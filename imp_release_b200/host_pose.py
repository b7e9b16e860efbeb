"""Host-side relative-pose estimation from matches (the reference's eval/pose_estimation.py:92-115,
called from eval/eval_imp.py:167-173 and eval/matching.py:84).  The north star keeps this step ON THE HOST: it is OpenCV
RANSAC, not part of the GPU hot path.  It lives here only so that the evaluation loop of this repo
(``evaluate_pairs`` below: LatencyMatcher + PoseOverlap) can be run and measured on a box that has no reference tree; with
the reference present, pass its own ``estimate_pose`` instead -- the loop only needs a callable.

``estimate_pose`` follows the reference's contract: ``None`` for fewer than 5 matches or a degenerate essential matrix,
else ``(E, R, t, inlier_mask)`` with the rotation / translation picked among the four decompositions of E by the
cheirality test on the RANSAC inliers.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional

import numpy as np


def estimate_pose(kpts0, kpts1, K0, K1, norm_thresh, conf=0.99999, method=None, mask=None):
    import cv2
    if len(kpts0) < 5:
        return None
    method = cv2.RANSAC if method is None else method
    E, emask = cv2.findEssentialMat(points1=np.ascontiguousarray(kpts0, dtype=np.float64),
                                    points2=np.ascontiguousarray(kpts1, dtype=np.float64), cameraMatrix1=K0,
                                    cameraMatrix2=K1, distCoeffs1=None, distCoeffs2=None, threshold=norm_thresh, prob=conf,
                                    mask=mask, method=method)
    if E is None or E.shape != (3, 3):
        return None
    inl = emask.ravel() > 0
    K = (K0 + K1) / 2.0
    n0 = (np.asarray(kpts0, dtype=np.float64)[inl] - K[[0, 1], [2, 2]]) / K[[0, 1], [0, 1]]
    n1 = (np.asarray(kpts1, dtype=np.float64)[inl] - K[[0, 1], [2, 2]]) / K[[0, 1], [0, 1]]
    R1, R2, t = cv2.decomposeEssentialMat(E)
    t = t.reshape(3)
    P0 = np.eye(3, 4)
    best = None
    for R, tt in ((R1, t), (R2, t), (R1, -t), (R2, -t)):
        P = np.hstack([R, tt.reshape(3, 1)])
        X = cv2.triangulatePoints(P0, P, n0.T.copy(), n1.T.copy())
        good = (X[2] * X[3]) > 0
        X = X / X[3]
        good &= X[2] < 1000
        Y = P @ X
        good &= (Y[2] > 0) & (Y[2] < 1000)
        if best is None or good.sum() > best[2].sum():
            best = (R, tt, good)
    out = np.zeros(len(kpts0), dtype=bool)
    out[inl] = best[2]
    return E, best[0], best[1], out


def pose_from_matches(indices0: np.ndarray, mscores0: Optional[np.ndarray], kpts0: np.ndarray, kpts1: np.ndarray, K0, K1,
                      norm_thresh: float = 1.0, pose_fn: Callable = estimate_pose):
    """What eval/eval_imp.py:160-173 does with one pair's matches: gather the matched keypoints, estimate the pose."""
    valid = indices0 > -1
    return pose_fn(kpts0[valid], kpts1[indices0[valid]], K0, K1, norm_thresh)


def evaluate_pairs(matcher, pairs: Iterable[Dict], pose_fn: Callable = estimate_pose, workers: int = 4,
                   max_pending: int = 16, norm_thresh: float = 1.0) -> List:
    """Evaluation loop with the GPU and the host overlapped.  ``matcher`` is a LatencyMatcher (several pairs in flight on
    the GPU); every pair's matches go device -> pinned host asynchronously and its pose is estimated by a worker thread
    while the GPU already matches the following pairs (PoseOverlap).  ``pairs`` yields the reference's feed dicts
    (eval/eval_imp.py:59-78: tensors + 'pts0_cpu', 'pts1_cpu', 'K0', 'K1').  Returns the pose results in order."""
    import torch
    from .pose_overlap import PoseOverlap
    futs = []
    with PoseOverlap(workers=workers, max_pending=max_pending) as po:
        tickets = []
        for d in pairs:
            tickets.append((matcher.submit(d), d))
            if len(tickets) >= len(matcher.slots):             # keep `slots` pairs in flight on the GPU
                tk, dd = tickets.pop(0)
                out = matcher.result(tk)
                futs.append(po.submit(out['indices0'][-1][0], out['mscores0'][-1][0], pose_from_matches, dd['pts0_cpu'],
                                      dd['pts1_cpu'], dd['K0'], dd['K1'], norm_thresh, pose_fn))
        for tk, dd in tickets:
            out = matcher.result(tk)
            futs.append(po.submit(out['indices0'][-1][0], out['mscores0'][-1][0], pose_from_matches, dd['pts0_cpu'],
                                  dd['pts1_cpu'], dd['K0'], dd['K1'], norm_thresh, pose_fn))
        res = [f.result() for f in futs]
    return res
