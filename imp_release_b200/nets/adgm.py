"""AdaGMN ("EIMP") -- DGNNS-style matcher with adaptive pooling of keypoints (nets/adgm.py:15-635).

Batched ``produce_matches`` (nets/adgm.py:327-526): queries are never dropped; from the first update iteration on
the KEYS / VALUES of every attention call are restricted to the kept ids of their image.  The reference does that
with dense {0,1} masks over the [N, N] score matrix; here the kept K / V rows are physically compacted (gather
kernel) and the attention kernel simply sees fewer keys -- masked probabilities are exactly 0 in the reference, so
the two are equivalent.  The keep sets, per-sample Sinkhorn sizes and match scatter all stay on the device (one host read of the largest kept count per
pruning round, none per
iteration); only the final 'scores' slice needs the last sample's sizes.
"""
from __future__ import annotations

import torch

from .. import ops
from ..engine import D
from ..ops import Planes
from .gm import GM, AttentionStash
from .layers import SHARING_LAYERS, normalize_keypoints  # noqa: F401


class AdaGMN(GM):
    _sharing = SHARING_LAYERS

    def __init__(self, config={}):
        self.pool_sizes = [0, 0] * 2 + [0, 0, 0, 0] * 21
        self.sharing_layers = SHARING_LAYERS
        super().__init__(config={**config, **{'pool_sizes': self.pool_sizes}})
        self.n_min_tokens = self.config['n_min_tokens']
        self.with_ada = True
        self.first_it_to_update = 2

    # ------------------------------------------------------------------ batched path
    def _received(self, st):
        """Attention received per (kept) key: a_self[img], a_cross[img] indexed by key position of image img."""
        ws = st.ws
        dev = ws.H.device
        a_self = torch.empty(ws.n_img, ws.Np, dtype=torch.float32, device=dev)
        a_cross = torch.empty(ws.n_img, ws.Np, dtype=torch.float32, device=dev)
        eng = self.engine()
        eng.received_attention(st, 'self', a_self)
        eng.received_attention(st, 'cross', a_cross)      # rows already belong to the key image
        return a_self, a_cross

    @ops.on_model_device
    def produce_matches(self, data, p=0.2, mscore_th=0.1, uncertainty_ratio=1., **kwargs):
        desc0, desc1 = data['descriptors0'], data['descriptors1']
        nk0, nk1 = self._norm_kpts(data)
        st = self._begin(desc0, desc1, nk0, nk1, data['scores0'], data['scores1'])
        eng = self.engine()
        ws = st.ws
        B, N0, N1, Np = st.B, st.N0, st.N1, ws.Np
        dev = desc0.device
        nI = self.config['n_layers']
        thresh = float(mscore_th * uncertainty_ratio)
        all_i0, all_m0 = [], []
        last_sk = None
        yc = None
        S0, S1 = N0, N1            # capacity of the kept-subset problems (distance GEMM, Sinkhorn), see the update below
        for ni in range(nI):
            eng.layer(st, 2 * ni)
            eng.layer(st, 2 * ni + 1)
            update = ni >= self.first_it_to_update and self.sharing_layers[2 * ni]
            if st.key_ids is None:
                _, i0, _, m0, _, sk = self._score(st, ni, p, keep_scores=False, want_mass=update,
                                                  write_scores=(ni == nI - 1))
                n0s = n1s = None
            else:
                # Sinkhorn on the kept subsets: gather the projected descriptors of the kept tokens
                eng.project(st, ni)
                if yc is None:
                    yc = Planes.empty((2 * B, Np, D), dev)
                y3h, y3l = ws.Y.hi.view(2 * B, Np, D), ws.Y.lo.view(2 * B, Np, D)
                ops.gather_rows(y3h, st.key_ids, st.key_cnt, yc.hi, Np)
                ops.gather_rows(y3l, st.key_ids, st.key_cnt, yc.lo, Np)
                ldd = (S1 + 7) // 8 * 8
                dist = self._dist_buffer(B, N0, (N1 + 7) // 8 * 8, dev).view(-1)[:B * S0 * ldd].view(B, S0, ldd)
                eng.distance(st, Planes(yc.hi.view(-1, D), yc.lo.view(-1, D)), S0, S1, dist, ldd)
                n0s, n1s = st.key_cnt[:B], st.key_cnt[B:]
                _, i0c, _, m0c, _, sk = self._score_from_dist(dist, ldd, B, S0, S1, p, False, want_mass=update,
                                                              n0s=n0s, n1s=n1s, write_scores=(ni == nI - 1))
                i0 = torch.full((B, N0), -1, dtype=torch.int64, device=dev)
                m0 = torch.zeros(B, N0, dtype=torch.float32, device=dev)
                ops.scatter_matches(i0c, m0c, st.key_ids[:B], st.key_ids[B:], n0s, i0, m0)
            all_i0.append(i0); all_m0.append(m0)
            last_sk = sk
            if update:
                a_self, a_cross = self._received(st)
                mass = torch.zeros(2 * B, Np, dtype=torch.float32, device=dev)
                mass[:B, :sk.row_mass.shape[1]] = sk.row_mass
                mass[B:, :sk.col_mass.shape[1]] = sk.col_mass
                if st.key_ids is None:
                    ids_in = torch.arange(Np, dtype=torch.int32, device=dev).repeat(2 * B, 1).contiguous()
                    cnt_in = st.n_tok
                else:
                    ids_in, cnt_in = st.key_ids, st.key_cnt
                ids_out, cnt_out, _ = ops.pool_select(mass, a_self, a_cross, ids_in, cnt_in, thresh, self.n_min_tokens)
                st.key_ids, st.key_cnt = ids_out, cnt_out
                # ONE host read per pruning round (3 per forward).  Kept sets only shrink, so the largest kept set of this
                # round bounds every later distance / Sinkhorn problem: sizing them by it (instead of N with per-sample
                # counts) halves the score GEMM and lets the Sinkhorn ring buffers hold full-width rows again -- with
                # N-wide slots and ~1100-wide rows only half the bytes were in flight (measured at batch 128: Sinkhorn
                # 101 of 197 ms).  Rounded up to 64 so that consecutive batches reuse the cached workspaces.
                kmax = cnt_out.view(2, B).max(dim=1).values.tolist()
                S0, S1 = min(N0, (kmax[0] + 63) // 64 * 64), min(N1, (kmax[1] + 63) // 64 * 64)
        if last_sk is not None and st.key_ids is not None and nI > self.first_it_to_update and sk is last_sk and n0s is not None:
            c0, c1 = int(n0s[B - 1]), int(n1s[B - 1])
            scores = [last_sk.P[B - 1:B, :c0 + 1, :c1 + 1].clone()]      # the workspace is cached and reused: hand out a copy
        elif last_sk is not None:
            scores = [(last_sk.scores()[B - 1:B] if nI > self.first_it_to_update else last_sk.scores()).clone()]
        else:
            scores = [None]
        zero, one = torch.zeros([], device=dev), torch.ones([], device=dev)
        self._kept = (st.key_cnt, st.key_ids)
        return {'scores': scores, 'indices0': all_i0, 'mscores0': all_m0, 'acc_corr': [zero], 'acc_incorr': [zero],
                'total_acc_corr': [one], 'total_acc_incorr': [one]}

    def run(self, data):
        """nets/adgm.py:607-635."""
        out = self.produce_matches_test(
            data={'descriptors0': data['desc1'], 'descriptors1': data['desc2'],
                  'norm_keypoints0': data['x1'][:, :, :2], 'norm_keypoints1': data['x2'][:, :, :2],
                  'scores0': data['x1'][:, :, -1], 'scores1': data['x2'][:, :, -1]},
            p=self.config['match_threshold'])
        indices0 = out['indices0'][-1][0]
        index0 = torch.where(indices0 >= 0)[0]
        return {'index0': index0, 'index1': indices0[index0]}

    def produce_matches_test(self, data, p=0.2, only_last=False, **kwargs):
        return self.produce_matches(data=data, p=p)

    # ------------------------------------------------------------------ B = 1 iterative path
    @ops.on_model_device
    def pool(self, pred_score, prob00, prob01, prob11, prob10, mscore_th=0.1, uncertainty_ratio=1.0, n_min_tokens=256):
        """AdaGMN.pool (nets/adgm.py:552-605): which keypoints of each image to keep; the caller compacts
        (eval/matching.py:166-174).  ``prob*`` are the opaque AttentionStash handles of this model."""
        st = self._st
        if st is None or st.B != 1:
            raise RuntimeError('pool() follows forward_one_layer() on a single pair (eval/matching.py:254)')
        for h in (prob00, prob01, prob11, prob10):
            if not isinstance(h, AttentionStash):
                raise TypeError('pool() expects the attention handles returned by this model (model.self_prob0, ...)')
        N0, N1 = pred_score.shape[1] - 1, pred_score.shape[2] - 1
        if (N0, N1) != (st.N0, st.N1):
            raise RuntimeError(f'pool(): pred_score is {N0} x {N1} but the stashed attention belongs to {st.N0} x {st.N1} keypoints')
        dev = pred_score.device
        Np = st.ws.Np
        sk = self._last_sk
        if sk is not None and sk.P.data_ptr() == pred_score.data_ptr() and sk.row_mass is not None:
            row_mass, col_mass = sk.row_mass, sk.col_mass
        else:
            _, _, _, row_mass, col_mass = ops.score_argmax(pred_score, N0, N1, want_mass=True)
        a_self, a_cross = self._received(st)
        mass = torch.zeros(2, Np, dtype=torch.float32, device=dev)
        mass[0, :N0] = row_mass[0]
        mass[1, :N1] = col_mass[0]
        ids_in = torch.arange(Np, dtype=torch.int32, device=dev).repeat(2, 1).contiguous()
        # the reference compares the AUGMENTED sizes (pred_score.shape, nets/adgm.py:554-555,567-574) with n_min_tokens
        ids_out, cnt_out, changed = ops.pool_select(mass, a_self, a_cross, ids_in, st.n_tok,
                                                    float(mscore_th * uncertainty_ratio), max(n_min_tokens - 1, 0))
        cnt, chg = cnt_out.tolist(), changed.tolist()          # data-dependent output shapes: one host sync
        ids0 = ids_out[0, :cnt[0]].long() if chg[0] else None
        ids1 = ids_out[1, :cnt[1]].long() if chg[1] else None
        return ids0, ids1
