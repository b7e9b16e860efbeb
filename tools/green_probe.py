"""Can the HBM-bound Sinkhorn phase run NEXT TO the tensor-bound GNN layers?  Splits the GPU's SMs into two green contexts
(CUDA driver API), runs the Sinkhorn scoring on one partition and attention + the MLP GEMMs on the other, alone and
concurrently.  usage: python tools/green_probe.py [sm counts for the Sinkhorn partition ...]"""
import json
import sys

import torch

sys.path.insert(0, '.')
from cuda.bindings import driver as drv  # noqa: E402

from imp_release_b200 import ops  # noqa: E402


def ck(res):
    err, *rest = res
    if int(err) != 0:
        raise RuntimeError(f'CUDA driver error {err}')
    return rest[0] if len(rest) == 1 else rest


torch.cuda.init()
torch.zeros(1, device='cuda')
dev = ck(drv.cuDeviceGet(0))
full = ck(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print('device SMs', full.sm.smCount, flush=True)


def partition(n_s):
    groups, nb, rem = ck(drv.cuDevSmResourceSplitByCount(1, full, 0, n_s))
    out = []
    for r in (groups[0], rem):
        desc = ck(drv.cuDevResourceGenerateDesc([r], 1))
        g = ck(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        s = ck(drv.cuGreenCtxStreamCreate(g, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        out.append((r.sm.smCount, g, torch.cuda.ExternalStream(int(s))))
    return out


B, N = 64, 2000
dist = torch.randn(B, N, N, device='cuda') * 3
bs = torch.tensor(1.0, device='cuda')
ws = ops.SinkhornWorkspace(B, N, N, 'cuda', storage='fp32')
n_img = 2 * B
g = torch.Generator('cuda').manual_seed(0)
qkv = (torch.randn(n_img, N, 768, device='cuda', generator=g) * 1.2).half()
base = qkv.data_ptr()
lse = torch.zeros(n_img, 4, N, device='cuda')
out = ops.Planes.empty((n_img, N, 256), 'cuda')
M = n_img * N
xa = ops.split_planes(torch.randn(M, 512, device='cuda', generator=g))
wb = ops.split_planes(torch.randn(512, 512, device='cuda', generator=g) * 0.05)
H = torch.zeros(M, 512, device='cuda')


def sink():
    ops.sinkhorn(dist, N, bs, 20, ws, write_scores=False)


def tensor_work():  # one attention call + two MLP0-sized GEMMs ~ 1.6 ms of tensor-bound work
    ops.attention(base, base + 512, base + 1024, n_img=n_img, src_offset=0, Nq_max=N, Nk_max=N, nq=None, nk=None, shared=False,
                  lse=lse, out=out, q_row_stride=768, kv_row_stride=768)
    for _ in range(2):
        ops.gemm(xa, wb, M=M, N=512, K1=512, a_row_stride=512, b_row_stride=512, out_mode=ops.OUT_F32, out0=H, out_row_stride=512)


def timed(fn, stream, n=5, sms=0):
    ops.set_option(ops.OPT_SM_LIMIT, sms)
    with torch.cuda.stream(stream):
        for _ in range(2):
            fn()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
main = torch.cuda.current_stream()
res['full_device'] = {'sinkhorn_ms': timed(sink, main), 'tensor_ms': timed(tensor_work, main)}
print(res['full_device'], flush=True)
for n_s in [int(v) for v in sys.argv[1:]] or [32, 48, 64]:
    try:
        (cs, gs, ss), (cg, gg, sg) = partition(n_s)
    except Exception as e:  # noqa: BLE001
        print('partition', n_s, 'failed:', e, flush=True)
        continue
    row = {'sms_sinkhorn': cs, 'sms_tensor': cg}
    row['sinkhorn_alone_ms'] = timed(sink, ss, sms=cs)
    row['tensor_alone_ms'] = timed(tensor_work, sg, sms=cg)
    # concurrently: k rounds of (2 x tensor_work) next to k Sinkhorn scorings
    torch.cuda.synchronize()
    e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    k = 4
    e0.record(ss)
    e2.record(sg)
    for _ in range(k):  # interleave the submissions so that both partitions have work queued
        ops.set_option(ops.OPT_SM_LIMIT, cs)
        with torch.cuda.stream(ss):
            sink()
        ops.set_option(ops.OPT_SM_LIMIT, cg)
        with torch.cuda.stream(sg):
            tensor_work()
            tensor_work()
    ops.set_option(ops.OPT_SM_LIMIT, 0)
    e1.record(ss)
    e3.record(sg)
    torch.cuda.synchronize()
    row['concurrent_sinkhorn_ms'] = e0.elapsed_time(e1) / k
    row['concurrent_tensor_ms'] = e2.elapsed_time(e3) / (2 * k)
    res[f'split_{n_s}'] = row
    print(row, flush=True)
json.dump(res, open('gpurun_out/green_probe.json', 'w'), indent=1)
