import sys, torch
sys.path.insert(0, '.')
from imp_release_b200 import ops
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
T = 256000
def planes(r, c):
    return ops.Planes(torch.randn(r, c, device='cuda').half(), (torch.randn(r, c, device='cuda') * 1e-3).half())
X, A, Hn = planes(T, 256), planes(T, 256), planes(T, 512)
for name, N, K1, K2, mode in (('qkv n768 k256 f16', 768, 256, 0, ops.OUT_F16), ('mlp0 n512 k512(2seg) f32', 512, 256, 256, ops.OUT_F32),
                              ('mlp1 n256 k512 resid', 256, 512, 0, ops.OUT_SPLIT_RESID), ('proj n256 k256 split', 256, 256, 0, ops.OUT_SPLIT)):
    W = planes(N, K1 + K2); bias = torch.randn(N, device='cuda')
    a = Hn if K1 == 512 else X
    if mode == ops.OUT_F16: o0 = torch.empty(T, N, device='cuda', dtype=torch.float16); o1 = None
    elif mode == ops.OUT_F32: o0 = torch.empty(T, N, device='cuda'); o1 = None
    else: o0 = torch.empty(T, N, device='cuda', dtype=torch.float16); o1 = torch.empty_like(o0)
    res = ops.Planes(o0, o1) if mode == ops.OUT_SPLIT_RESID else None
    ms = t(lambda: ops.gemm(a, W, M=T, N=N, K1=K1, K2=K2, a2=(A if K2 else None), a_row_stride=K1, a2_row_stride=K2, b_row_stride=K1 + K2,
                            bias=bias, out_mode=mode, out0=o0, out1=o1, out_row_stride=N, res=res))
    ms1 = t(lambda: ops.gemm(a, W, M=T, N=N, K1=K1, K2=K2, a2=(A if K2 else None), a_row_stride=K1, a2_row_stride=K2, b_row_stride=K1 + K2,
                            bias=bias, out_mode=mode, out0=o0, out1=o1, out_row_stride=N, res=res, nsplit=1))
    print(f'    nsplit=1: {ms1:.3f} ms')
    fl = 2.0 * T * N * (K1 + K2)
    print(f'{name:28s} {ms:.3f} ms  useful {fl/ms/1e9:7.1f} TF/s  (x3 MMA {3*fl/ms/1e9:7.1f})')
B, N = 64, 2000
Y = planes(2 * B * N, 256); dist = torch.empty(B, N, N, device='cuda')
ms = t(lambda: ops.gemm(Y, Y, M=N, N=N, K1=256, batch=B, a_row_stride=256, a_batch_stride=N * 256, b_row_stride=256, b_batch_stride=N * 256,
                        b_batched=True, b_offset=B * N * 256, alpha=1 / 16, out_mode=ops.OUT_F32, out0=dist, out_row_stride=N, out_batch_stride=N * N))
fl = 2.0 * B * N * N * 256
print(f'{"dist batched n2000 k256":28s} {ms:.3f} ms  useful {fl/ms/1e9:7.1f} TF/s  (x3 MMA {3*fl/ms/1e9:7.1f})')
