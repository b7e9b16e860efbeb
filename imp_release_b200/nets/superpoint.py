"""SuperPoint front-end with the reference's Python API (nets/superpoint.py:97-235), executed by the sm_100a kernels of
libimp_b200.so (csrc/superpoint.cu): image -> keypoints, scores, 256-d descriptors, i.e. the inputs of the matcher.

Same class name, constructor config, ``state_dict`` keys / shapes (``conv1a.weight`` ... ``convDb.bias``, so the published
``superpoint_v1.pth`` loads unchanged), ``forward`` / ``extract`` signatures and return layout as the reference.  There is no
CPU / PyTorch fallback: parameters live in ``nn.Conv2d`` containers that are never called.

Kernel schedule (one launch each unless noted): conv1a (SIMT, C_in = 1) -> 7 tcgen05 implicit-GEMM 3 x 3 convolutions, the 3
max-pools fused into their epilogues -> convPa / convDa (3 x 3, tcgen05) -> convPb / convDb (1 x 1 = the matcher's split-precision GEMM) ->
65-way softmax + depth-to-space -> simple_nms (5 launches) -> ordered compaction + top-k (4 launches per image) -> descriptor
L2 normalisation -> bilinear sampling + L2 normalisation.  Activations are NHWC fp16 hi/lo planes.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from .. import ops
from ..ops import Planes


class SuperPoint(nn.Module):
    default_config = {                 # nets/superpoint.py:106-112
        'descriptor_dim': 256,
        'nms_radius': 4,
        'keypoint_threshold': 0.0025,
        'max_keypoints': -1,
        'remove_borders': 4,
    }

    # (name, C_in, C_out, kernel): shared encoder, detector head (Pa, Pb), descriptor head (Da, Db)
    LAYERS = (('conv1a', 1, 64, 3), ('conv1b', 64, 64, 3), ('conv2a', 64, 64, 3), ('conv2b', 64, 64, 3), ('conv3a', 64, 128, 3),
              ('conv3b', 128, 128, 3), ('conv4a', 128, 128, 3), ('conv4b', 128, 128, 3), ('convPa', 128, 256, 3),
              ('convPb', 256, 65, 1), ('convDa', 128, 256, 3), ('convDb', 256, 256, 1))

    def __init__(self, config):
        super().__init__()
        self.config = {**self.default_config, **config}
        if self.config['descriptor_dim'] != 256:
            raise NotImplementedError('the B200 kernels are specialised for 256-d descriptors')
        # parameter containers only -- same names, shapes and registration order as nets/superpoint.py:122-143; never called
        for name, cin, cout, k in self.LAYERS:
            setattr(self, name, nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=k // 2))
        # the reference requires config['weight_path'] (nets/superpoint.py:146-147); None / absent = keep the initialisation
        # (B200-side allowance: pretrained weights are not always at hand, e.g. in the parity tests)
        path = self.config.get('weight_path', None)
        if path is not None:
            self.load_state_dict(torch.load(str(path), map_location='cpu'))
            print('Loaded SuperPoint model')
        mk = self.config['max_keypoints']
        if mk == 0 or mk < -1:
            raise ValueError('"max_keypoints" must be positive or "-1"')
        self._packed = None
        self._packed_key = None
        self._sel_ws: Dict[tuple, ops.SpSelectWorkspace] = {}

    # ------------------------------------------------------------------ weight packing
    def _apply(self, fn, *a, **kw):
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._packed = None
        return super().load_state_dict(*a, **kw)

    def _weights(self):
        w0 = self.conv1a.weight
        key = (w0.device, w0.data_ptr(), w0._version)
        if self._packed is None or self._packed_key != key:
            if not w0.is_cuda:
                raise ops._lib.ImpLibraryError('move the model to a CUDA device first (net.cuda()); the B200 path has no CPU '
                                               'implementation')
            P = {}
            f = lambda t: t.detach().float().contiguous()
            P['conv1a'] = (f(self.conv1a.weight).reshape(64, 9).contiguous(), f(self.conv1a.bias))
            for n in ('conv1b', 'conv2a', 'conv2b', 'conv3a', 'conv3b', 'conv4a', 'conv4b', 'convPa', 'convDa'):
                m = getattr(self, n)
                w = f(m.weight).permute(0, 2, 3, 1).reshape(m.weight.shape[0], -1).contiguous()   # [Cout, (ky, kx, ci)]
                P[n] = (ops.split_planes(w), f(m.bias))
            wp = torch.zeros(128, 256, device=w0.device)
            wp[:65] = f(self.convPb.weight)[:, :, 0, 0]
            bp = torch.zeros(128, device=w0.device)
            bp[:65] = f(self.convPb.bias)
            P['convPb'] = (ops.split_planes(wp), bp)
            P['convDb'] = (ops.split_planes(f(self.convDb.weight)[:, :, 0, 0].contiguous()), f(self.convDb.bias))
            self._packed, self._packed_key = P, key
        return self._packed

    # ------------------------------------------------------------------ dense part
    def _dense(self, image: torch.Tensor):
        """image [B, 1, H, W] -> (scores [B, 8 Hc, 8 Wc] before NMS, normalised descriptor map [B, Hc, Wc, 256])."""
        if image.dim() != 4 or image.shape[1] != 1:
            raise ValueError('SuperPoint expects a [B, 1, H, W] grayscale image')
        ops._require_cuda(image)
        P = self._weights()
        dev = image.device
        B, _, H, W = image.shape
        if H < 8 or W < 8:
            raise ValueError('image smaller than one 8 x 8 cell')
        img = image.reshape(B, H, W).float().contiguous()

        def conv(x: Planes, name: str, pool: bool = False) -> Planes:
            w, b = P[name]
            b_, h_, w_, _ = x.hi.shape
            if pool:                       # nn.MaxPool2d(2, 2) fused into the convolution's epilogue
                h_, w_ = h_ // 2, w_ // 2
            out = Planes(torch.empty((b_, h_, w_, w.hi.shape[0]), dtype=torch.float16, device=dev),
                         torch.empty((b_, h_, w_, w.hi.shape[0]), dtype=torch.float16, device=dev))
            return ops.sp_conv3x3(x, w, b, out, relu=True, pool=pool)

        x = Planes(torch.empty((B, H, W, 64), dtype=torch.float16, device=dev), torch.empty((B, H, W, 64), dtype=torch.float16, device=dev))
        x = ops.sp_conv1a(img, P['conv1a'][0], P['conv1a'][1], x)
        x = conv(x, 'conv1b', pool=True)
        x = conv(conv(x, 'conv2a'), 'conv2b', pool=True)
        x = conv(conv(x, 'conv3a'), 'conv3b', pool=True)
        x = conv(conv(x, 'conv4a'), 'conv4b')
        Hc, Wc = x.hi.shape[1], x.hi.shape[2]
        M = B * Hc * Wc
        # detector head
        cPa = conv(x, 'convPa')
        logits = torch.empty(M, 128, dtype=torch.float32, device=dev)
        wp, bp = P['convPb']
        ops.gemm(Planes(cPa.hi.view(M, 256), cPa.lo.view(M, 256)), wp, M=M, N=128, K1=256, a_row_stride=256, b_row_stride=256,
                 bias=bp, out_mode=ops.OUT_F32, out0=logits, out_row_stride=128)
        scores = torch.empty(B, Hc * 8, Wc * 8, dtype=torch.float32, device=dev)
        ops.sp_scores(logits, scores, B, Hc, Wc)
        # descriptor head
        cDa = conv(x, 'convDa')
        dmap = torch.empty(M, 256, dtype=torch.float32, device=dev)
        wd, bd = P['convDb']
        ops.gemm(Planes(cDa.hi.view(M, 256), cDa.lo.view(M, 256)), wd, M=M, N=256, K1=256, a_row_stride=256, b_row_stride=256,
                 bias=bd, out_mode=ops.OUT_F32, out0=dmap, out_row_stride=256)
        ops.sp_l2norm_rows(dmap)
        return scores, dmap.view(B, Hc, Wc, 256)

    @ops.on_model_device
    def extract(self, data):
        """Dense scores [B, H, W] (no NMS) and descriptors [B, 256, Hc, Wc] (nets/superpoint.py:154-183)."""
        scores, dmap = self._dense(data['image'])
        return scores, dmap.permute(0, 3, 1, 2)

    @ops.on_model_device
    def detect_padded(self, image: torch.Tensor):
        """B200-side entry point WITHOUT any host synchronisation: fixed-capacity outputs in the matcher's input layout,
        {'keypoints': [B, K, 2] (x, y), 'scores': [B, K], 'descriptors': [B, K, 256], 'n_keypoints': [B] int32 (device)},
        K = config['max_keypoints'] > 0, rows behind the n real keypoints of an image zeroed.  Hand the tensors to
        ``produce_matches`` as ``keypoints0/scores0/descriptors0/n_keypoints0`` (imp_release_b200/pipeline.py): the whole
        image pair -> matches path then runs without the host ever waiting for the GPU (``forward`` has to read the keypoint
        count back, like the reference's torch.nonzero)."""
        K = self.config['max_keypoints']
        if K <= 0:
            raise ValueError('detect_padded needs a positive max_keypoints (the fixed output capacity)')
        scores, dmap = self._dense(image)
        B, Hs, Ws = scores.shape
        Hc, Wc = dmap.shape[1], dmap.shape[2]
        dev = scores.device
        mask = torch.empty(B, Hs, Ws, dtype=torch.uint8, device=dev)
        supp = torch.empty(B, Hs, Ws, dtype=torch.uint8, device=dev)
        ops.sp_nms(scores, mask, supp, self.config['nms_radius'])
        ws = self._select_ws(Hs, Ws, K, dev)
        kpts = torch.empty(B, K, 2, dtype=torch.float32, device=dev)
        ksc = torch.empty(B, K, dtype=torch.float32, device=dev)
        desc = torch.empty(B, K, 256, dtype=torch.float32, device=dev)
        cnt = torch.empty(B, dtype=torch.int32, device=dev)
        for b in range(B):
            ops.sp_select(scores[b], mask[b], ws, self.config['keypoint_threshold'], self.config['remove_borders'], K)
            kpts[b].copy_(ws.kpts[:K])          # stream-ordered copies out of the shared selection workspace
            ksc[b].copy_(ws.kscores[:K])
            cnt[b:b + 1].copy_(ws.n_out)
            ops.sp_sample_descriptors(dmap[b], kpts[b], cnt[b:b + 1], desc[b], Hc, Wc, K)
        return {'keypoints': kpts, 'scores': ksc, 'descriptors': desc, 'n_keypoints': cnt}

    def _select_ws(self, Hs, Ws, mk, dev):
        # one workspace per (shape, stream): calls on different streams may be in flight at the same time
        key = (Hs, Ws, mk, str(dev), torch.cuda.current_stream(dev).cuda_stream)
        ws = self._sel_ws.get(key)
        if ws is None:
            if len(self._sel_ws) > 16:
                self._sel_ws.clear()
            ws = self._sel_ws[key] = ops.SpSelectWorkspace(Hs, Ws, mk, dev)
        return ws

    @ops.on_model_device
    def forward(self, data):
        """{'image': [B, 1, H, W]} -> {'keypoints': [[K, 2] (x, y)], 'scores': [[K]], 'descriptors': [[256, K]]}
        (nets/superpoint.py:185-235)."""
        scores, dmap = self._dense(data['image'])
        B, Hs, Ws = scores.shape
        Hc, Wc = dmap.shape[1], dmap.shape[2]
        dev = scores.device
        mask = torch.empty(B, Hs, Ws, dtype=torch.uint8, device=dev)
        supp = torch.empty(B, Hs, Ws, dtype=torch.uint8, device=dev)
        ops.sp_nms(scores, mask, supp, self.config['nms_radius'])
        mk = self.config['max_keypoints']
        ws = self._select_ws(Hs, Ws, mk, dev)
        keypoints: List[torch.Tensor] = []
        kscores: List[torch.Tensor] = []
        descriptors: List[torch.Tensor] = []
        for b in range(B):
            ops.sp_select(scores[b], mask[b], ws, self.config['keypoint_threshold'], self.config['remove_borders'], mk)
            n = int(ws.n_out.item())          # the reference synchronises here too (torch.nonzero)
            kp = ws.kpts[:n].clone()
            sc = ws.kscores[:n].clone()
            desc = torch.empty(max(n, 1), 256, dtype=torch.float32, device=dev)
            if n > 0:
                ops.sp_sample_descriptors(dmap[b], kp, None, desc, Hc, Wc, n)
            keypoints.append(kp)
            kscores.append(sc)
            descriptors.append(desc[:n].t())
        return {'keypoints': keypoints, 'scores': kscores, 'descriptors': descriptors}
