"""GM -- plain (non-sharing) matcher with the reference's Python API (nets/gm.py:16-364), executed by the sm_100a
kernels of libimp_b200.so.  DGNNS (IMP) and AdaGMN (EIMP) derive from it exactly like in the reference."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..engine import D, Engine, PackedModel, RunState
from ..ops import Planes, SinkhornWorkspace
from .layers import GNNParams, KeypointEncoderParams, normalize_keypoints, _conv  # noqa: F401


class AttentionStash:
    """Opaque stand-in for the reference's ``prob`` tensors (``model.self_prob0`` ... , eval/matching.py:185-188).
    The B200 path never materialises the [B,4,N,M] attention map; the handle names the stashed Q / K / row-LSE from
    which ``pool`` recomputes the only thing callers ever derive from it: the attention received per source token."""

    def __init__(self, name: str, side: int, token: int):
        self.name, self.side, self.token = name, side, token

    def __repr__(self):
        return f'AttentionStash({self.name}, queries of image {self.side})'


class _ScoreHolder:
    """What the dual-softmax scorer hands to the EIMP loop in place of a SinkhornWorkspace."""

    def __init__(self, scores, row_mass, col_mass):
        self.P, self.row_mass, self.col_mass = scores, row_mass, col_mass

    def scores(self):
        return self.P


class GM(nn.Module):
    default_config = {                 # nets/gm.py:30-44
        'descriptor_dim': 256,
        'weights': 'indoor',
        'keypoint_encoder': [32, 64, 128, 256],
        'GNN_layers': ['self', 'cross'] * 9,
        'sinkhorn_iterations': 20,
        'match_threshold': 0.2,
        'with_pose': False,
        'n_layers': 9,
        'n_min_tokens': 256,
        'with_sinkhorn': True,
        'ac_fn': 'relu',
        'norm_fn': 'bn',
    }
    _sharing: Optional[List[bool]] = None      # subclasses: SHARING_LAYERS

    def __init__(self, config):
        super().__init__()
        self.config = {**self.default_config, **config}
        if self.config['descriptor_dim'] != D or list(self.config['keypoint_encoder']) != [32, 64, 128, 256]:
            raise NotImplementedError('the B200 kernels are specialised for 256-d descriptors (SuperPoint) and the '
                                      '[32,64,128,256] keypoint encoder')
        if self.config['norm_fn'] != 'in' or self.config['ac_fn'] != 'relu':
            raise NotImplementedError("only norm_fn='in', ac_fn='relu' (the eval configuration, eval/eval_imp.py:259-270) "
                                      'is implemented')
        self.n_layers = self.config['n_layers']
        self.with_sinkhorn = self.config['with_sinkhorn']
        self.match_threshold = self.config['match_threshold']
        self.sinkhorn_iterations = self.config['sinkhorn_iterations']
        self.kenc = KeypointEncoderParams(D, self.config['keypoint_encoder'])
        self.gnn = GNNParams(D, self.config['GNN_layers'], self._sharing)
        self.final_proj = nn.ModuleList([_conv(D, D) for _ in range(self.n_layers)])
        self.register_parameter('bin_score', nn.Parameter(torch.tensor(1.)))
        # B200-specific knob (not a reference config key): 'fp16' = single fp16 attention operands (default, fastest);
        # 'high' = 3-product split for QK^T and PV (fp32-level attention; use when the attention is sharply peaked)
        self.attention_precision = self.config.get('attention_precision', 'fp16')
        if self.attention_precision not in ('fp16', 'high'):
            raise ValueError("attention_precision must be 'fp16' or 'high'")
        # second B200-specific knob: how softmax(M) is stored for the Sinkhorn iteration sweeps of big batches
        # ('fp32' | 'fp24' | 'fp16', include/imp_b200.h IMP_SK_STORE_*); None = library default (IMP_SK_STORAGE or 'fp32')
        self.sinkhorn_storage = self.config.get('sinkhorn_storage', None)
        if self.sinkhorn_storage not in (None, 'fp32', 'fp24', 'fp16'):
            raise ValueError("sinkhorn_storage must be 'fp32', 'fp24' or 'fp16'")
        self.self_prob0 = self.self_prob1 = self.cross_prob0 = self.cross_prob1 = None
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self._st: Optional[RunState] = None          # per-layer API state (stateful like the reference)
        self._token = 0
        self._sk_cache: Dict[tuple, SinkhornWorkspace] = {}
        self._last_sk = None

    # ------------------------------------------------------------------ engine / packing
    def _apply(self, fn, *a, **kw):            # .cuda() / .to() / .float(): parameters move -> repack lazily
        self._engine = None
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):       # new weights -> repack lazily
        self._engine = None
        return super().load_state_dict(*a, **kw)

    def invalidate_packed_weights(self):
        """Call after modifying parameters in place (load_state_dict / .to() are tracked automatically)."""
        self._engine = None

    def engine(self) -> Engine:
        # cheap staleness check on the hot path; the full per-parameter scan only runs when (re)packing
        key = (self.bin_score.device, self.bin_score.data_ptr(), self.kenc.encoder[0].weight._version)
        if self._engine is None or key != self._engine_key:
            params = list(self.parameters())
            if not params[0].is_cuda:
                raise ops._lib.ImpLibraryError('move the model to a CUDA device first (net.cuda()); the B200 path has '
                                               'no CPU implementation')
            sd = {k: v for k, v in self.state_dict().items()}
            n_gnn = len(self.gnn.layers)
            self._engine = Engine(PackedModel(sd, self.n_layers, self.gnn.sharing_layers), self.gnn.names,
                                  high_precision_attention=(self.attention_precision == 'high'),
                                  stash_lo=getattr(self, 'with_ada', False))
            assert n_gnn >= 2 * self.n_layers
            self._engine_key = key
        return self._engine

    def _sinkhorn_ws(self, B, N0, N1, device, want_mass=False, fresh=False) -> SinkhornWorkspace:
        res = getattr(self, 'sinkhorn_resident', True)
        if fresh:
            return SinkhornWorkspace(B, N0, N1, device, want_mass, storage=self.sinkhorn_storage, resident=res)
        key = (B, N0, N1, str(device), want_mass)
        if key not in self._sk_cache:
            if len(self._sk_cache) > 8:
                self._sk_cache.clear()
            self._sk_cache[key] = SinkhornWorkspace(B, N0, N1, device, want_mass, storage=self.sinkhorn_storage, resident=res)
        return self._sk_cache[key]

    def replica(self):
        """A second handle on the same weights (parameters and packed kernel weights are shared, nothing is copied) with
        its OWN workspaces and per-call state, so that several pairs can be in flight on different CUDA streams
        (imp_release_b200.graphed.LatencyMatcher).  The reference model is not re-entrant either (stateful attributes)."""
        import copy
        eng = self.engine()
        r = copy.copy(self)                       # shares _parameters / _modules
        r._engine = Engine(eng.pk, eng.names, high_precision_attention=eng.hp, stash_lo=eng.stash_lo)
        r._engine_key = self._engine_key
        r._st, r._io, r._last_sk = None, None, None
        r._sk_cache = {}
        r.__dict__.pop('_n_tok_cache', None)
        r.__dict__.pop('_dist', None)
        r.__dict__.pop('_dist_key', None)
        r.__dict__.pop('_side', None)
        return r

    # ------------------------------------------------------------------ batched entry points
    @ops.on_model_device
    def forward(self, data, mode=0):
        if self.training:
            raise NotImplementedError('training (forward_train) is out of scope of the B200 inference path')
        if mode == 0:
            return self.produce_matches(data=data)
        return self.run(data=data)

    @staticmethod
    def _empty_result(kpts0, kpts1):          # nets/gm.py:154-163
        shape0, shape1 = kpts0.shape[:-1], kpts1.shape[:-1]
        return {
            'matches0': kpts0.new_full(shape0, -1, dtype=torch.int)[0],
            'matches1': kpts1.new_full(shape1, -1, dtype=torch.int)[0],
            'matching_scores0': kpts0.new_zeros(shape0)[0],
            'matching_scores1': kpts1.new_zeros(shape1)[0],
            'skip_train': True,
        }

    @staticmethod
    def _norm_kpts(data):
        if 'norm_keypoints0' in data.keys() and 'norm_keypoints1' in data.keys():
            return data['norm_keypoints0'], data['norm_keypoints1']
        if 'image0' in data.keys() and 'image1' in data.keys():
            return (normalize_keypoints(data['keypoints0'], data['image0'].shape),
                    normalize_keypoints(data['keypoints1'], data['image1'].shape))
        raise ValueError('Require image shape for keypoint coordinate normalization')

    @staticmethod
    def _bucket(B: int, n: int) -> int:
        """Token capacity of the workspace for n keypoints per image.  Small problems (the one-pair-per-call evaluation,
        eval/eval_imp.py:155-173, where every pair has its own keypoint counts) round up to a multiple of 128 so that
        workspaces -- and captured CUDA graphs -- are shared by all pairs of a bucket; the kernels mask by the per-image
        counts.  Big batches keep the exact size (no padded rows in the GEMMs)."""
        return (n + 127) // 128 * 128 if 2 * B * n <= 32768 else n

    @staticmethod
    def _counts(data, dev):
        """B200 extension of the data dict: optional 'n_keypoints0' / 'n_keypoints1' ([B] integer tensors) declare that
        the keypoint tensors are zero-padded and only the first n_keypoints*[b] entries of pair b are real (ragged pairs
        in one batch, static-shape graph replay).  The reference has no such key; without it nothing changes."""
        c0, c1 = data.get('n_keypoints0'), data.get('n_keypoints1')
        if c0 is None or c1 is None:
            return None
        return torch.cat([c0.reshape(-1), c1.reshape(-1)]).to(device=dev, dtype=torch.int32)

    def _begin(self, desc0, desc1, nk0, nk1, sc0, sc1, counts: Optional[torch.Tensor] = None) -> RunState:
        """Stack both images token-major, run the keypoint encoder, x = desc + enc (nets/gms.py:158-172)."""
        eng = self.engine()
        self._io = None            # the workspace state is about to be overwritten
        B, N0, N1 = desc0.shape[0], desc0.shape[1], desc1.shape[1]
        dev = desc0.device
        Np = self._bucket(B, max(N0, N1))
        ws = eng.workspace(2 * B, Np, dev)
        n_tok = counts if counts is not None else self._n_tok(B, N0, N1, dev)
        f32 = dict(dtype=torch.float32, device=dev)
        if N0 == N1 == Np:
            nk = torch.cat([nk0, nk1], 0).float().contiguous()
            sc = torch.cat([sc0, sc1], 0).float().contiguous()
        else:
            nk = torch.zeros(2 * B, Np, 2, **f32)
            sc = torch.zeros(2 * B, Np, **f32)
            nk[:B, :N0], nk[B:, :N1] = nk0, nk1
            sc[:B, :N0], sc[B:, :N1] = sc0, sc1
        eng.encode_keypoints(ws, nk, sc, n_tok, ws.tok_f32)
        if N0 == N1 == Np and desc0.dtype == torch.float32 and desc0.is_contiguous() and desc1.is_contiguous():
            # descriptors go straight from the caller's tensors into the hi/lo planes (no stacked fp32 copy)
            h = B * Np
            for side, d in ((0, desc0), (1, desc1)):
                ops.split_planes(d.view(-1, D), out=Planes(ws.X.hi[side * h:(side + 1) * h], ws.X.lo[side * h:(side + 1) * h]),
                                 addend=ws.tok_f32[side * h:(side + 1) * h])
        else:
            desc = torch.zeros(2 * B, Np, D, **f32)
            desc[:B, :N0], desc[B:, :N1] = desc0, desc1
            ops.split_planes(desc.view(-1, D), out=ws.X, addend=ws.tok_f32)
        st = RunState(ws, B, N0, N1, n_tok)
        st.ragged = counts is not None
        return st

    def _n_tok(self, B, N0, N1, dev) -> torch.Tensor:
        """Per-image token counts [2B] (cached: no host-to-device copy on the hot path, CUDA-graph capturable)."""
        key = (B, N0, N1, str(dev))
        cache = self.__dict__.setdefault('_n_tok_cache', {})
        if key not in cache:
            if len(cache) > 64:
                cache.clear()
            cache[key] = torch.tensor([N0] * B + [N1] * B, dtype=torch.int32, device=dev)
        return cache[key]

    def _score(self, st: RunState, ni: int, p: float, keep_scores: bool, want_mass: bool = False, write_scores=None):
        """final_proj -> dist -> Sinkhorn / dual-softmax -> mutual matches (nets/gm.py:290-320) on the full sets."""
        eng = self.engine()
        B, N0, N1 = st.B, st.N0, st.N1
        dev = st.ws.H.device
        eng.project(st, ni)
        ldd = (N1 + 7) // 8 * 8
        dist = self._dist_buffer(B, N0, ldd, dev)
        eng.distance(st, st.ws.Y, N0, N1, dist, ldd)
        n0s, n1s = (st.n_tok[:B], st.n_tok[B:]) if st.ragged else (None, None)
        return self._score_from_dist(dist, ldd, B, N0, N1, p, keep_scores, want_mass, n0s=n0s, n1s=n1s,
                                     write_scores=write_scores)

    def _dist_buffer(self, B, N0, ldd, dev):
        key = (B, N0, ldd, str(dev))
        if getattr(self, '_dist_key', None) != key:
            self._dist = torch.zeros(B, N0, ldd, dtype=torch.float32, device=dev)
            self._dist_key = key
        return self._dist

    def _score_from_dist(self, dist, ldd, B, N0, N1, p, keep_scores, want_mass=False, n0s=None, n1s=None,
                         dist_batch_stride=None, write_scores=None):
        """keep_scores: the score matrix is handed to the caller (fresh buffer, final scaling written back);
        otherwise the Sinkhorn only produces arg-max / masses and skips the write-back sweep."""
        dev = dist.device
        if write_scores is None:
            write_scores = keep_scores
        if self.with_sinkhorn:
            sk = self._sinkhorn_ws(B, N0, N1, dev, want_mass, fresh=keep_scores)
            ops.sinkhorn(dist, ldd, self.bin_score.data, self.sinkhorn_iterations, sk, n0s=n0s, n1s=n1s,
                         dist_batch_stride=dist_batch_stride, write_scores=write_scores)
            i0, i1, m0, m1 = ops.matches(sk.row_max, sk.row_arg, sk.col_key, p, N0, N1, B, n0s=n0s, n1s=n1s)
            self._last_sk = sk
            return sk.scores(), i0, i1, m0, m1, sk
        # with_sinkhorn=False (eval_imp.py --use_dual_softmax): same contract, masses from the arg-max pass
        scores = ops.dual_softmax(dist, ldd, self.bin_score.data, N0, N1, B, n0s=n0s, n1s=n1s,
                                  dist_batch_stride=dist_batch_stride)
        am = ops.score_argmax(scores, N0, N1, want_mass=want_mass, n0s=n0s, n1s=n1s)
        i0, i1, m0, m1 = ops.matches(am[0], am[1], am[2], p, N0, N1, B, n0s=n0s, n1s=n1s)
        holder = _ScoreHolder(scores, am[3] if want_mass else None, am[4] if want_mass else None)
        return scores, i0, i1, m0, m1, holder

    @ops.on_model_device
    def produce_matches(self, data, p=0.2, only_last=False, **kwargs):
        """GM.produce_matches (nets/gm.py:145-247): all GNN layers, then scoring of every iteration (or the last)."""
        desc0, desc1 = data['descriptors0'], data['descriptors1']
        kpts0, kpts1 = data['keypoints0'], data['keypoints1']
        if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:
            return self._empty_result(kpts0, kpts1)
        nk0, nk1 = self._norm_kpts(data)
        st = self._begin(desc0, desc1, nk0, nk1, data['scores0'], data['scores1'], self._counts(data, desc0.device))
        eng = self.engine()
        nI = len(eng.names) // 2
        all_scores, all_i0, all_m0 = [], [], []
        for ni in range(nI):
            eng.layer(st, 2 * ni)
            eng.layer(st, 2 * ni + 1)
            if only_last and ni != nI - 1:
                continue
            lid = (self.n_layers - 1) if only_last else ni
            sc, i0, _, m0, _, _ = self._score(st, lid, p, keep_scores=True)
            all_scores.append(sc); all_i0.append(i0); all_m0.append(m0)
        dev = desc0.device
        zero, one = torch.zeros([], device=dev), torch.ones([], device=dev)
        return {'scores': all_scores, 'indices0': all_i0, 'mscores0': all_m0, 'acc_corr': [zero],
                'acc_incorr': [zero], 'total_acc_corr': [one], 'total_acc_incorr': [one]}

    def produce_matches_test(self, data, p=0.2, only_last=False, **kwargs):
        return self.produce_matches(data=data, p=p, only_last=only_last, kwargs=kwargs)

    def run(self, data):
        """GM.run (nets/gm.py:322-364): mode=1 adapter returning the last score matrix."""
        out = self.produce_matches(data={'descriptors0': data['desc1'], 'descriptors1': data['desc2'],
                                         'keypoints0': data['x1'][:, :, :2], 'keypoints1': data['x2'][:, :, :2],
                                         'norm_keypoints0': data['x1'][:, :, :2], 'norm_keypoints1': data['x2'][:, :, :2],
                                         'scores0': data['x1'][:, :, -1], 'scores1': data['x2'][:, :, -1]},
                                   p=self.match_threshold, only_last=True)
        return {'p': out['scores'][-1]}

    # ------------------------------------------------------------------ per-layer API (eval/matching.py)
    @ops.on_model_device
    def encode_keypoint(self, norm_kpts0, norm_kpts1, scores0, scores1):
        """nets/gm.py:287-288 -> (enc0 [B,256,N0], enc1 [B,256,N1]) fp32, channels-first like the reference."""
        eng = self.engine()
        B, N0, N1 = norm_kpts0.shape[0], norm_kpts0.shape[1], norm_kpts1.shape[1]
        dev = norm_kpts0.device
        Np = self._bucket(B, max(N0, N1))
        ws = eng.workspace(2 * B, Np, dev)
        nk = norm_kpts0.new_zeros(2 * B, Np, 2, dtype=torch.float32)
        sc = norm_kpts0.new_zeros(2 * B, Np, dtype=torch.float32)
        nk[:B, :N0], nk[B:, :N1] = norm_kpts0, norm_kpts1
        sc[:B, :N0], sc[B:, :N1] = scores0, scores1
        n_tok = self._n_tok(B, N0, N1, dev)
        enc = torch.empty(2 * B * Np, D, dtype=torch.float32, device=dev)
        eng.encode_keypoints(ws, nk, sc, n_tok, enc)
        enc = enc.view(2 * B, Np, D)
        return enc[:B, :N0].transpose(1, 2), enc[B:, :N1].transpose(1, 2)

    def _load_state(self, desc0, desc1) -> RunState:
        """Channels-first caller tensors -> token-major planes in the workspace (boundary conversion).  If the caller
        hands back exactly the tensors the previous call returned (eval/matching.py:51-58 does), the planes already in
        the workspace are the full-precision state and the conversion is skipped."""
        io = getattr(self, '_io', None)
        if (io is not None and desc0 is io[0] and desc1 is io[1] and desc0._version == io[2] and desc1._version == io[3]
                and self._st is not None and self._engine is not None):
            return self._st
        eng = self.engine()
        B, N0, N1 = desc0.shape[0], desc0.shape[2], desc1.shape[2]
        dev = desc0.device
        Np = self._bucket(B, max(N0, N1))
        ws = eng.workspace(2 * B, Np, dev)
        x = desc0.new_zeros(2 * B, Np, D, dtype=torch.float32)
        x[:B, :N0] = desc0.transpose(1, 2)
        x[B:, :N1] = desc1.transpose(1, 2)
        ops.split_planes(x.view(-1, D), out=ws.X)
        st = self._st
        if st is None or st.ws is not ws or st.B != B or st.N0 != N0 or st.N1 != N1:
            st = RunState(ws, B, N0, N1, self._n_tok(B, N0, N1, dev))
            self._st = st
        return st

    def _export_state(self, st: RunState):
        x = st.ws.X.float().view(2 * st.B, st.ws.Np, D)
        o0, o1 = x[:st.B, :st.N0].transpose(1, 2), x[st.B:, :st.N1].transpose(1, 2)
        self._io = (o0, o1, o0._version, o1._version)
        return o0, o1

    @ops.on_model_device
    def forward_one_layer(self, desc0, desc1, M0, M1, layer_i):
        """nets/gms.py:260-282 / nets/adgm.py:528-550: one self or cross layer on both images; stateful (the stashed
        attention of a non-sharing layer is consumed by the sharing layer of the next iteration)."""
        st = self._load_state(desc0, desc1)
        self.engine().layer(st, layer_i)
        self._token += 1
        name = self.engine().names[layer_i]
        if name == 'cross':
            self.cross_prob1 = AttentionStash('cross', 0, self._token)   # queries image 0 -> keys image 1 (prob10)
            self.cross_prob0 = AttentionStash('cross', 1, self._token)   # queries image 1 -> keys image 0 (prob01)
        else:
            self.self_prob0 = AttentionStash('self', 0, self._token)
            self.self_prob1 = AttentionStash('self', 1, self._token)
        return self._export_state(st)

    @ops.on_model_device
    def compute_distance(self, desc0, desc1, layer_id=-1):
        """nets/gm.py:290-295 -> dist [B,N0,N1] (a view of a padded buffer)."""
        st = self._load_state(desc0, desc1)
        eng = self.engine()
        lid = layer_id % self.n_layers
        eng.project(st, lid)
        ldd = (st.N1 + 7) // 8 * 8
        dist = torch.zeros(st.B, st.N0, ldd, dtype=torch.float32, device=desc0.device)
        eng.distance(st, st.ws.Y, st.N0, st.N1, dist, ldd)
        return dist[:, :, :st.N1]

    @ops.on_model_device
    def compute_score(self, dist, dustbin, iteration):
        """nets/gm.py:297-303 -> scores [B,N0+1,N1+1] (a view of a padded buffer; a real torch.Tensor)."""
        if dist.stride(2) != 1 or dist.stride(1) % 4 != 0:
            d = dist.new_zeros(dist.shape[0], dist.shape[1], (dist.shape[2] + 3) // 4 * 4)
            d[:, :, :dist.shape[2]] = dist
            dist = d[:, :, :dist.shape[2]]
        B, N0, N1 = dist.shape
        bin_t = dustbin.data if isinstance(dustbin, torch.Tensor) else torch.tensor(float(dustbin), device=dist.device)
        if self.with_sinkhorn:
            sk = self._sinkhorn_ws(B, N0, N1, dist.device, want_mass=True, fresh=True)
            ops.sinkhorn(dist, dist.stride(1), bin_t.float(), iteration, sk, dist_batch_stride=dist.stride(0))
            self._last_sk = sk
            return sk.scores()
        self._last_sk = None
        return ops.dual_softmax(dist, dist.stride(1), bin_t.float(), N0, N1, B, dist_batch_stride=dist.stride(0))

    @ops.on_model_device
    def compute_matches(self, scores, p=0.2):
        """nets/gm.py:305-320."""
        B, N0, N1 = scores.shape[0], scores.shape[1] - 1, scores.shape[2] - 1
        sk = self._last_sk
        if sk is not None and sk.P.data_ptr() == scores.data_ptr() and (sk.batch, sk.N0max, sk.N1max) == (B, N0, N1):
            return ops.matches(sk.row_max, sk.row_arg, sk.col_key, p, N0, N1, B)      # arg-max fused in Sinkhorn
        rmx, rarg, ckey = ops.score_argmax(scores, N0, N1)
        return ops.matches(rmx, rarg, ckey, p, N0, N1, B)

    def pool(self, **kwargs):
        return None, None
