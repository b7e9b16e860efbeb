"""Fused instance norm (GEMM-epilogue statistics + apply) vs the stand-alone slab kernel: timing and agreement."""
import sys
import torch
sys.path.insert(0, '.')
from imp_release_b200 import ops


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n_img, Np, C = 128, 2000, 512
T = n_img * Np
g = torch.Generator('cuda').manual_seed(0)
def planes(r, c, scale=1.0):
    return ops.Planes((torch.randn(r, c, device='cuda', generator=g) * scale).half(), (torch.randn(r, c, device='cuda', generator=g) * 1e-3).half())
X, A = planes(T, 256), planes(T, 256)
W = planes(512, 512, 0.05)
bias = torch.randn(512, device='cuda', generator=g)
ns = torch.randint(1500, Np + 1, (n_img,), device='cuda', generator=g, dtype=torch.int32)
ns[0] = Np
H = torch.zeros(T, C, device='cuda')
st = ops.InstNormStats(n_img, Np, C, 'cuda')
out_a, out_b = ops.Planes.empty((T, C), 'cuda'), ops.Planes.empty((T, C), 'cuda')
gem = lambda stats: ops.gemm(X, W, M=T, N=512, K1=256, K2=256, a2=A, a_row_stride=256, a2_row_stride=256, b_row_stride=512, bias=bias,
                             out_mode=ops.OUT_F32, out0=H, out_row_stride=512, stats=stats, ns=ns, Np=Np)
print('gemm plain  %.4f ms' % t(lambda: gem(None)))
print('gemm +stats %.4f ms' % t(lambda: gem(st)))
print('instnorm slab   %.4f ms' % t(lambda: ops.instnorm_relu(H, batch=n_img, Nmax=Np, C_=C, ns=ns, out=out_a)))
print('instnorm apply  %.4f ms' % t(lambda: ops.instnorm_apply(H, st, batch=n_img, Nmax=Np, C_=C, ns=ns, out=out_b)))
a, b = out_a.float().view(n_img, Np, C), out_b.float().view(n_img, Np, C)
err = 0.0
for i in range(n_img):
    err = max(err, float((a[i, :int(ns[i])] - b[i, :int(ns[i])]).abs().max()))
print('max |slab - fused| over valid rows: %.3e' % err)
h = H.view(n_img, Np, C)[3, :int(ns[3])].double()
ref = torch.relu((h - h.mean(0)) / torch.sqrt(h.var(0, unbiased=False) + 1e-3))
print('fused vs fp64 reference (image 3): %.3e ; slab vs fp64: %.3e' % (float((b[3, :int(ns[3])].double() - ref).abs().max()), float((a[3, :int(ns[3])].double() - ref).abs().max())))

# ---- consumer side (K5): second MLP GEMM with the normalisation in its A-operand path vs apply pass + plain GEMM
W3 = planes(256, 512, 0.05); b3 = torch.randn(256, device='cuda', generator=g)
Xr = planes(T, 256)
def mlp1_plain(o):
    ops.instnorm_apply(H, st, batch=n_img, Nmax=Np, C_=C, ns=ns, out=out_b)
    ops.gemm(out_b, W3, M=T, N=256, K1=512, a_row_stride=512, b_row_stride=512, bias=b3, out_mode=ops.OUT_SPLIT_RESID,
             out0=o.hi, out1=o.lo, out_row_stride=256, res=Xr)
def mlp1_fused(o):
    ops.instnorm_apply(H, st, batch=n_img, Nmax=Np, C_=C, ns=ns, out=None)
    ops.gemm(out_b, W3, M=T, N=256, K1=512, a_row_stride=512, b_row_stride=512, bias=b3, out_mode=ops.OUT_SPLIT_RESID,
             out0=o.hi, out1=o.lo, out_row_stride=256, res=Xr, a_f32=H, a_stats=st, Np=Np)
gem(st)
o1, o2 = ops.Planes.empty((T, 256), 'cuda'), ops.Planes.empty((T, 256), 'cuda')
mlp1_plain(o1); mlp1_fused(o2); torch.cuda.synchronize()
print('MLP1 fused-A vs apply+GEMM: max |diff| = %.3e (max |out| %.2f)' % (float((o1.float() - o2.float()).abs().max()), float(o1.float().abs().max())))
print('apply + gemm      %.4f ms' % t(lambda: mlp1_plain(o1)))
print('finalize + gemm(A=norm(H)) %.4f ms' % t(lambda: mlp1_fused(o2)))
