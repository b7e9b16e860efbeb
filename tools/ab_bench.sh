#!/bin/bash
# A/B of optional paths inside ONE box: bench main part only.  usage: ab_bench.sh
run() { env "$@" python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']
print('$*', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], {n: k[n]['avg_ms'] for n in ('sinkhorn','gemm_n512_k512','gemm_n256_k512','gemm_n768_k256','instnorm_apply_c512') if n in k})"; }
run A=default
run IMP_FUSED_NORM_A=0
run A=default
