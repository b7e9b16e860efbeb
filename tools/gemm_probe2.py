"""CTA-pair GEMM (variant 2) vs the single-CTA kernel (variant 0): agreement and timing on the shapes of the bench step."""
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


g = torch.Generator('cuda').manual_seed(0)
def planes(r, c, scale=1.0):
    return ops.Planes((torch.randn(r, c, device='cuda', generator=g) * scale).half(), (torch.randn(r, c, device='cuda', generator=g) * 1e-3).half())
small = len(sys.argv) > 1 and sys.argv[1] == 'small'
T = 2 * 1000 if small else 256000
X, A, Hn = planes(T, 256), planes(T, 256), planes(T, 512)
cases = (('qkv n768 k256 f16', 768, 256, 0, ops.OUT_F16), ('mlp0 n512 k512(2seg) f32', 512, 256, 256, ops.OUT_F32),
         ('mlp1 n256 k512 resid', 256, 512, 0, ops.OUT_SPLIT_RESID), ('proj n256 k256 split', 256, 256, 0, ops.OUT_SPLIT))
for name, N, K1, K2, mode in cases:
    W = planes(N, K1 + K2, 0.05); bias = torch.randn(N, device='cuda', generator=g)
    a = Hn if K1 == 512 else X
    outs = {}
    for variant in (0, 2):
        ops.set_option(ops.OPT_GEMM_VARIANT, variant)
        if mode == ops.OUT_F16: o0 = torch.zeros(T, N, device='cuda', dtype=torch.float16); o1 = None
        elif mode == ops.OUT_F32: o0 = torch.zeros(T, N, device='cuda'); o1 = None
        else: o0 = torch.zeros(T, N, device='cuda', dtype=torch.float16); o1 = torch.zeros_like(o0)
        res = planes(T, N) if mode == ops.OUT_SPLIT_RESID else None
        fn = lambda: ops.gemm(a, W, M=T, N=N, K1=K1, K2=K2, a2=(A if K2 else None), a_row_stride=K1, a2_row_stride=K2, b_row_stride=K1 + K2,
                              bias=bias, out_mode=mode, out0=o0, out1=o1, out_row_stride=N, res=res)
        if res is not None:
            torch.manual_seed(1); res.hi.copy_(torch.randn(T, N, device='cuda').half()); res.lo.zero_()
        fn(); torch.cuda.synchronize()
        outs[variant] = (o0.float() + (o1.float() if o1 is not None else 0)).clone()
        ms = t(fn) if not small else 0.0
        fl = 2.0 * T * N * (K1 + K2)
        print(f'{name:28s} variant {variant}: {ms:.3f} ms  x3 MMA {3*fl/max(ms,1e-9)/1e9:7.1f} TF/s', flush=True)
    print('   max |v2 - v0| = %.3e   (max |v0| %.2f)' % (float((outs[2] - outs[0]).abs().max()), float(outs[0].abs().max())), flush=True)
B, N = (2, 300) if small else (64, 2000)
Y = planes(2 * B * N, 256, 0.5)
for variant in (0, 2):
    ops.set_option(ops.OPT_GEMM_VARIANT, variant)
    ld = (N + 7) // 8 * 8
    dist = torch.zeros(B, N, ld, device='cuda')
    fn = lambda: ops.gemm(Y, Y, M=N, N=N, K1=256, batch=B, a_row_stride=256, a_batch_stride=N * 256, b_row_stride=256, b_batch_stride=N * 256,
                          b_batched=True, b_offset=B * N * 256, alpha=1 / 16, out_mode=ops.OUT_F32, out0=dist, out_row_stride=ld, out_batch_stride=N * ld)
    fn(); torch.cuda.synchronize()
    if variant == 0: d0 = dist.clone()
    ms = t(fn) if not small else 0.0
    print(f'dist batched n{N} k256 variant {variant}: {ms:.3f} ms', flush=True)
print('   max |v2 - v0| = %.3e' % float((dist - d0).abs().max()))
ops.set_option(ops.OPT_GEMM_VARIANT, 0)
