"""GPU diagnostic: per-layer deviation of the B200 path from the fp32 oracle (not a test; prints a table)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import imp_oracle, synth
from imp_release_b200 import DGNNS, GM, AdaGMN
from imp_release_b200.engine import D

def main():
    N0, N1, B, nl = int(sys.argv[1]) if len(sys.argv) > 1 else 500, int(sys.argv[2]) if len(sys.argv) > 2 else 460, 2, 9
    cfg = dict(n_layers=nl, GNN_layers=['self', 'cross'] * nl, norm_fn='in', ac_fn='relu', sinkhorn_iterations=20,
               with_sinkhorn=True, descriptor_dim=256)
    sd = synth.make_state_dict('DGNNS', nl, seed=7)
    data = synth.make_pair_batch(seed=3, batch=B, n0=N0, n1=N1)
    orc = imp_oracle.Oracle('DGNNS', cfg, sd)
    net = DGNNS(cfg); net.load_state_dict(sd, strict=True); net = net.cuda().eval()
    gdata = {k: v.cuda() for k, v in data.items()}
    with torch.no_grad():
        x0, x1 = orc._prepare(data)
        nk0, nk1 = net._norm_kpts(gdata)
        st = net._begin(gdata['descriptors0'], gdata['descriptors1'], nk0, nk1, gdata['scores0'], gdata['scores1'])
        eng = net.engine()
        def cmp(tag):
            x = st.ws.X.float().view(2 * B, st.ws.Np, D).cpu()
            d0 = (x[:B, :N0] - x0).abs().max().item(); d1 = (x[B:, :N1] - x1).abs().max().item()
            print(f'{tag:12s} max|dx0| {d0:.3e} max|dx1| {d1:.3e}  (|x| max {x0.abs().max().item():.2f})')
        cmp('encode')
        p00 = p11 = p10 = p01 = None
        for ni in range(nl):
            d0, p00 = orc._layer(2 * ni, x0, x0, p00); d1, p11 = orc._layer(2 * ni, x1, x1, p11)
            x0, x1 = x0 + d0, x1 + d1
            eng.layer(st, 2 * ni); cmp(f'it{ni} self')
            d0, p10 = orc._layer(2 * ni + 1, x0, x1, p10); d1, p01 = orc._layer(2 * ni + 1, x1, x0, p01)
            x0, x1 = x0 + d0, x1 + d1
            eng.layer(st, 2 * ni + 1); cmp(f'it{ni} cross')
            sc, i0, i1, m0, m1, _ = net._score(st, ni, 0.2, keep_scores=False)
            rs, ri0, rm0 = orc._score_and_match(x0, x1, ni, 0.2)
            print(f'   score it{ni}: idx mismatch {(i0.cpu() != ri0).sum().item()} matches {(ri0 >= 0).sum().item()} '
                  f'max|dms| {(m0.cpu() - rm0).abs().max().item():.2e} max|dscore| {(sc.cpu() - rs).abs().max().item():.2e}')
    torch.cuda.synchronize()
    print('done')

main()
