#!/bin/bash
# attention kernel variants: correctness (kernel tests under the variant) + timing probe.  usage: gpu_attn.sh <test variant> <probe variants...>
V=${1:-6}; shift
IMP_ATTN_VARIANT=$V timeout 300 python -m pytest tests/test_kernels_gpu.py -k "attention_self_cross" -x -q 2>&1 | tail -5
timeout 300 python tools/attn_probe.py "$@" 2>&1 | tail -8
