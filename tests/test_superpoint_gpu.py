"""GPU: the B200 SuperPoint front-end (imp_release_b200/nets/superpoint.py, csrc/superpoint.cu) against the outputs of the
UNMODIFIED reference class (tests/golden/reference_superpoint.npz) and, layer by layer, against the CPU oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import superpoint_oracle as spo
from tests.golden.make_golden_superpoint import CASES, DEFAULT, probe_dirs

pytestmark = pytest.mark.gpu
DEV = 'cuda'
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_superpoint.npz'))


def _net(wseed, over):
    from imp_release_b200.nets.superpoint import SuperPoint
    net = SuperPoint({**over})
    net.load_state_dict(spo.make_state_dict(wseed), strict=True)
    return net.eval().cuda()


@pytest.mark.parametrize('Cin,Cout,H,W,B', [(64, 64, 24, 40, 1), (64, 128, 19, 37, 2), (128, 128, 8, 16, 1), (128, 256, 15, 20, 3)])
def test_conv3x3_kernel(Cin, Cout, H, W, B):
    from imp_release_b200 import ops
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(x.double(), w.double(), bias.double(), padding=1))
    xp = ops.split_planes(x.permute(0, 2, 3, 1).contiguous().to(DEV))
    wp = ops.split_planes(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().to(DEV))
    out = ops.Planes.empty((B, H, W, Cout), DEV)
    ops.sp_conv3x3(xp, wp, bias.to(DEV), out, relu=True)
    pooled = ops.Planes.empty((B, H // 2, W // 2, Cout), DEV)      # fused nn.MaxPool2d(2, 2): same planes as pooling afterwards
    ops.sp_conv3x3(xp, wp, bias.to(DEV), pooled, relu=True, pool=True)
    assert torch.equal(pooled.float().permute(0, 3, 1, 2), F.max_pool2d(out.float().permute(0, 3, 1, 2), 2, 2))
    got = out.float().permute(0, 3, 1, 2).double().cpu()
    # fp32-level, not fp32-exact: the tensor core adds every 16-deep partial product into the TMEM accumulator with truncation,
    # so the error grows with K = 9 C_in (measured 7e-6 relative at C_in = 128; an fp16 / TF32 convolution is at 5e-4)
    assert float((got - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


def test_maxpool_and_first_layer():
    from imp_release_b200 import ops
    g = torch.Generator().manual_seed(3)
    img = torch.rand(2, 1, 21, 35, generator=g)
    w = torch.randn(64, 1, 3, 3, generator=g) * 0.4
    b = torch.randn(64, generator=g) * 0.1
    ref = F.relu(F.conv2d(img, w, b, padding=1))
    out = ops.Planes.empty((2, 21, 35, 64), DEV)
    ops.sp_conv1a(img[:, 0].contiguous().to(DEV), w.reshape(64, 9).contiguous().to(DEV), b.to(DEV), out)
    assert float((out.float().permute(0, 3, 1, 2).cpu() - ref).abs().max()) < 2e-6
    pooled = ops.sp_maxpool2(out, ops.Planes.empty((2, 10, 17, 64), DEV))
    refp = F.max_pool2d(out.float().permute(0, 3, 1, 2), 2, 2)
    assert torch.equal(pooled.float().permute(0, 3, 1, 2), refp)


@pytest.mark.parametrize('radius', [4, 3, 0])
def test_nms_and_selection_exact(radius):
    """simple_nms + nonzero + border removal + top-k on a synthetic score map with plateaus and exact ties: index work,
    bit-exact against the oracle restatement."""
    from imp_release_b200 import ops
    g = torch.Generator().manual_seed(radius)
    H, W = 93, 131
    s = torch.rand(1, H, W, generator=g) * 0.02
    s[0, 10:14, 20:30] = 0.5            # plateau: every pixel of it is a maximum of its window
    s[0, 50, 60] = s[0, 50, 64] = 0.9   # two equal peaks inside one window
    s = (s * 4096).round() / 4096       # many exact ties
    ref = spo.nms(s, radius)
    sc = s.to(DEV)
    mask = torch.empty(1, H, W, dtype=torch.uint8, device=DEV)
    supp = torch.empty_like(mask)
    ops.sp_nms(sc, mask, supp, radius)
    assert torch.equal(torch.where(mask.bool(), sc, torch.zeros_like(sc)).cpu(), ref)
    for mk in (-1, 40, 5000, 100000):     # radix select (k <= 4096), global bitonic sort (radius 0: ~10 k candidates), all
        k_ref, s_ref = spo.detect(ref[0], 0.0025, 4, mk)
        ws = ops.SpSelectWorkspace(H, W, mk, DEV)
        ops.sp_select(sc[0], mask[0], ws, 0.0025, 4, mk)
        n = int(ws.n_out.item())
        assert n == len(k_ref)
        got_s = ws.kscores[:n].cpu()
        assert torch.equal(got_s, s_ref)                       # same scores in the same (descending / row-major) order
        got_k = ws.kpts[:n].cpu()
        if mk < 0 or mk >= int(ws.total.item()):
            assert torch.equal(got_k, k_ref)                   # row-major order is fully determined
        else:                                                  # among EQUAL scores torch.topk's choice / order is unspecified
            a, b = set(map(tuple, got_k.tolist())), set(map(tuple, k_ref.tolist()))
            for x, y in a ^ b:                                 # only candidates tied with the k-th score may differ
                assert float(ref[0, int(y), int(x)]) == float(s_ref[-1])
            for i in range(n):                                 # and every keypoint carries its own score
                assert float(ref[0, int(got_k[i, 1]), int(got_k[i, 0])]) == float(got_s[i])


@pytest.mark.parametrize('name', list(CASES))
def test_superpoint_matches_reference(name):
    wseed, iseed, H, W, B, over = CASES[name]
    net = _net(wseed, over)
    img = spo.make_image(iseed, H, W, B)
    with torch.no_grad():
        out = net({'image': img.to(DEV)})
        dense_scores, dmap = net.extract({'image': img.to(DEV)})
    assert float(dense_scores.double().sum()) == pytest.approx(float(G[f'{name}/dense_scores_sum']), rel=1e-6)
    assert np.abs(dense_scores[:, ::8, ::8].cpu().numpy() - G[f'{name}/dense_scores_8x']).max() < 2e-6
    for b in range(B):
        k = out['keypoints'][b].cpu().numpy().astype(np.int32)
        s = out['scores'][b].cpu().numpy()
        d = out['descriptors'][b].cpu()
        assert d.shape == (256, len(k))
        kr, sr = G[f'{name}/{b}/keypoints'], G[f'{name}/{b}/scores']
        assert len(k) == len(kr)
        # keypoint SET identical; order identical wherever the reference's scores are not within fp32 noise of each other
        assert sorted(map(tuple, k.tolist())) == sorted(map(tuple, kr.tolist()))
        if not np.array_equal(k, kr):
            swapped = np.nonzero((k != kr).any(1))[0]
            assert np.abs(sr[swapped][:, None] - sr[swapped][None, :]).min(1, initial=1.0, where=~np.eye(len(swapped), dtype=bool)).max() < 2e-6
        order = {tuple(p): i for i, p in enumerate(k.tolist())}
        perm = np.array([order[tuple(p)] for p in kr.tolist()])
        rel = float((np.abs(s[perm] - sr) / sr).max())
        print(f'{name}[{b}]: {len(k)} keypoints, max relative score error {rel:.2e}')
        assert rel < 3e-5
        dref_head = G[f'{name}/{b}/descriptors_head']
        assert np.abs(d[:, perm[:48]].numpy() - dref_head).max() < 2e-5
        assert np.abs((d.t() @ probe_dirs()).numpy()[perm] - G[f'{name}/{b}/descriptor_probes']).max() < 1e-4


def test_extractor_interface(tmp_path):
    """ExtractSuperpoint (components/extractors.py:50-89): image file -> kpt [N, 3], desc [N, 256] with unit-norm rows."""
    import cv2
    from imp_release_b200.extractors import ExtractSuperpoint
    img = (spo.make_image(5, 200, 260)[0, 0].numpy() * 255).astype(np.uint8)
    path = str(tmp_path / 'img.png')
    cv2.imwrite(path, img)
    wpath = str(tmp_path / 'w.pth')
    torch.save(spo.make_state_dict(11), wpath)
    ex = ExtractSuperpoint({'det_th': 0.005, 'num_kpt': 120, 'resize': [160], 'weight_path': wpath})
    kpt, desc = ex.run(path)
    assert kpt.shape == (120, 3) and desc.shape == (120, 256)
    assert np.abs(np.linalg.norm(desc, axis=1) - 1).max() < 1e-5
    assert (np.diff(kpt[:, 2]) <= 0).all()          # top-k: descending scores
    assert kpt[:, 0].max() < 260 and kpt[:, 1].max() < 200


def test_image_pair_pipeline_without_host_sync():
    """ImagePairMatcher (SuperPoint.detect_padded -> produce_matches with device-side keypoint counts) against the two-step
    path with exact-size tensors (SuperPoint.forward, host reads the counts, then the matcher)."""
    from imp_release_b200 import DGNNS
    from imp_release_b200.pipeline import ImagePairMatcher
    from oracle import synth
    img0 = spo.make_image(41, 240, 320).to(DEV)
    img1 = torch.roll(img0, shifts=(6, -9), dims=(2, 3)).contiguous()
    with torch.no_grad():            # capacity K a little above the number of keypoints the images really have
        cnt = [_net(11, {'max_keypoints': -1, 'keypoint_threshold': 0.03})({'image': im})['keypoints'][0].shape[0] for im in (img0, img1)]
    assert min(cnt) > 0
    K = max(cnt) + 37
    sp = _net(11, {'max_keypoints': K, 'keypoint_threshold': 0.03})
    nl = 3
    net = DGNNS(dict(n_layers=nl, GNN_layers=['self', 'cross'] * nl, norm_fn='in', ac_fn='relu', sinkhorn_iterations=20,
                     with_sinkhorn=True, descriptor_dim=256))
    net.load_state_dict(synth.make_state_dict('DGNNS', nl, seed=7))
    net = net.cuda().eval()
    with torch.no_grad():
        f0, f1 = sp({'image': img0}), sp({'image': img1})
        n0, n1 = f0['keypoints'][0].shape[0], f1['keypoints'][0].shape[0]
        assert 0 < n0 < K and 0 < n1 < K, (n0, n1)
        data = {'image0': img0, 'image1': img1}
        for i, f in ((0, f0), (1, f1)):
            data[f'keypoints{i}'] = f['keypoints'][0][None]
            data[f'scores{i}'] = f['scores'][0][None]
            data[f'descriptors{i}'] = f['descriptors'][0].t()[None].contiguous()
        ref = net.produce_matches(data, p=0.2, only_last=True)
        out = ImagePairMatcher(sp, net)(img0, img1)
    assert int(out['n_keypoints0']) == n0 and int(out['n_keypoints1']) == n1
    assert torch.equal(out['keypoints0'][:n0], f0['keypoints'][0]) and torch.equal(out['keypoints1'][:n1], f1['keypoints'][0])
    assert float(out['keypoints0'][n0:].abs().sum()) == 0.0
    assert torch.equal(out['indices0'][:n0], ref['indices0'][-1][0])
    assert float((out['mscores0'][:n0] - ref['mscores0'][-1][0]).abs().max()) < 1e-4
    assert int((out['indices0'][:n0] > -1).sum()) > 0          # the shifted image does produce matches
    assert bool((out['indices0'][n0:] == -1).all())
    # different image sizes go through two separate detections
    out2 = ImagePairMatcher(sp, net)(img0, img1[:, :, :200, :280].contiguous())
    assert out2['indices0'].shape == (K,)
    # the same pipeline as CUDA-graph replay, 3 pairs in flight on 2 slots, a second shape in between
    from imp_release_b200.pipeline import GraphedImagePairMatcher
    gm = GraphedImagePairMatcher(sp, net, slots=2)
    img2 = torch.roll(img0, shifts=(-3, 5), dims=(2, 3)).contiguous()
    tickets = [gm.submit(img0, img1), gm.submit(img0, img2), gm.submit(img0[:, :, :200], img1[:, :, :200]), gm.submit(img0, img1)]
    r = [gm.result(t) for t in tickets]
    torch.cuda.synchronize()
    for g_out in (r[0], r[3]):
        assert torch.equal(g_out['indices0'], out['indices0']) and torch.equal(g_out['keypoints1'], out['keypoints1'])
        assert float((g_out['mscores0'] - out['mscores0']).abs().max()) < 1e-4
    ref2 = ImagePairMatcher(sp, net)(img0, img2)
    assert torch.equal(r[1]['indices0'], ref2['indices0']) and int(r[1]['n_keypoints1']) == int(ref2['n_keypoints1'])
    assert gm.captures == 3            # slot 0: full-size shape + cropped shape, slot 1: full-size shape


def test_no_keypoints_is_an_empty_result():
    sp = _net(11, {'max_keypoints': 64, 'keypoint_threshold': 0.9})      # nothing passes
    img = spo.make_image(5, 64, 96).to(DEV)
    with torch.no_grad():
        out = sp({'image': img})
        pad = sp.detect_padded(img)
    assert out['keypoints'][0].shape == (0, 2) and out['scores'][0].shape == (0,) and out['descriptors'][0].shape == (256, 0)
    assert int(pad['n_keypoints'][0]) == 0 and float(pad['descriptors'].abs().sum()) == 0.0 and float(pad['keypoints'].abs().sum()) == 0.0
