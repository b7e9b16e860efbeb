"""DGNNS ("IMP") -- GM with attention sharing (nets/gms.py:15-317)."""
from __future__ import annotations

import torch

from .. import ops
from .gm import GM
from .layers import SHARING_LAYERS, normalize_keypoints  # noqa: F401


class DGNNS(GM):
    _sharing = SHARING_LAYERS

    def __init__(self, config={}):
        super().__init__(config=config)
        self.overlap_scoring = True      # run the Sinkhorn of iteration i concurrently with the layers of i+1

    def _side_stream(self, device):
        s = self.__dict__.get('_side')
        if s is None or s.device != device:
            s = torch.cuda.Stream(device=device)
            self.__dict__['_side'] = s
        return s

    @ops.on_model_device
    def produce_matches(self, data, p=0.2, only_last=False, **kwargs):
        """nets/gms.py:139-258: per iteration self layer, cross layer, then (every iteration or last only)
        final_proj -> Sinkhorn -> matches.  The reference also returns the four [B,4,N,M] attention maps per
        iteration ('prob00' ...); no caller reads them and the B200 path never materialises them, so those lists
        hold ``None`` placeholders (documented deviation)."""
        desc0, desc1 = data['descriptors0'], data['descriptors1']
        kpts0, kpts1 = data['keypoints0'], data['keypoints1']
        if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:
            return self._empty_result(kpts0, kpts1)
        nk0, nk1 = self._norm_kpts(data)
        st = self._begin(desc0, desc1, nk0, nk1, data['scores0'], data['scores1'], self._counts(data, desc0.device))
        eng = self.engine()
        nI = self.config['n_layers']
        all_i0, all_m0 = [], []
        overlap = self.overlap_scoring and not only_last and not torch.cuda.is_current_stream_capturing()
        if not overlap:
            for ni in range(nI):
                eng.layer(st, 2 * ni)
                eng.layer(st, 2 * ni + 1)
                if only_last and ni != nI - 1:
                    continue
                _, i0, _, m0, _, _ = self._score(st, ni, p, keep_scores=False)
                all_i0.append(i0); all_m0.append(m0)
        else:
            # Scoring of iteration ni (score GEMM, Sinkhorn, matches: HBM-bound) runs on a side stream while the main
            # stream already computes the GNN layers of iteration ni+1 (tensor / MUFU-bound).  Dependencies: the side
            # stream needs the projected descriptors Y(ni); the main stream must not overwrite Y before the score GEMM
            # of the previous iteration has read it.
            main = torch.cuda.current_stream()
            side = self._side_stream(desc0.device)
            dev = desc0.device
            B, N0, N1 = st.B, st.N0, st.N1
            ldd = (N1 + 7) // 8 * 8
            dist = self._dist_buffer(B, N0, ldd, dev)
            dist_done = None
            side.wait_stream(main)
            for ni in range(nI):
                eng.layer(st, 2 * ni)
                eng.layer(st, 2 * ni + 1)
                if dist_done is not None:
                    main.wait_event(dist_done)
                eng.project(st, ni)
                y_ready = torch.cuda.Event()
                y_ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(y_ready)
                    eng.distance(st, st.ws.Y, N0, N1, dist, ldd)
                    dist_done = torch.cuda.Event()
                    dist_done.record(side)
                    n0s, n1s = (st.n_tok[:B], st.n_tok[B:]) if st.ragged else (None, None)
                    _, i0, _, m0, _, _ = self._score_from_dist(dist, ldd, B, N0, N1, p, False, n0s=n0s, n1s=n1s)
                all_i0.append(i0); all_m0.append(m0)
            main.wait_stream(side)
            for t in all_i0 + all_m0:
                t.record_stream(main)
        none = [None] * nI
        return {'indices0': all_i0, 'mscores0': all_m0, 'prob00': list(none), 'prob01': list(none),
                'prob11': list(none), 'prob10': list(none)}

    def run(self, data):
        """nets/gms.py:284-314."""
        out = self.produce_matches(
            data={'descriptors0': data['desc1'], 'descriptors1': data['desc2'],
                  'keypoints0': data['x1'][:, :, :2], 'keypoints1': data['x2'][:, :, :2],
                  'norm_keypoints0': data['x1'][:, :, :2], 'norm_keypoints1': data['x2'][:, :, :2],
                  'scores0': data['x1'][:, :, -1], 'scores1': data['x2'][:, :, -1]},
            p=self.config['match_threshold'], only_last=True)
        indices0 = out['indices0'][-1][0]
        index0 = torch.where(indices0 >= 0)[0]
        return {'index0': index0, 'index1': indices0[index0]}
