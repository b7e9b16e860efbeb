import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=['default', 'streaming'])
def sk_path(request):
    """Model-level parity tests run twice: with the library's default Sinkhorn dispatch (small problems -> the
    shared-memory-resident kernel) and with that kernel switched off, so that the SAME reference fixtures also pin the
    streaming kernels (csrc/sinkhorn_q.cu) that the full-size batches of bench.py run."""
    from imp_release_b200 import ops
    ops.set_sinkhorn_resident(request.param != 'streaming')
    yield request.param
    ops.set_sinkhorn_resident(True)
