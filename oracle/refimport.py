"""Import the UNMODIFIED reference classes from /root/reference for oracle validation.
TEST INFRASTRUCTURE; works only in the build container (the GPU box has no /root/reference),
callers must check ``available()`` first.  Nothing is copied out of the reference tree.

Two shims, both documented in SURVEY.md 8(c):
  1. ``nets.gm.sink_algorithm`` hard-codes ``device='cuda'`` (nets/layers.py:41-44); it is replaced
     by a wrapper that builds the padded matrix and the marginals on ``M.device`` and then calls
     the reference's own, unmodified ``sinkhorn`` (nets/layers.py:27-35) for the iteration.
  2. ``matplotlib`` is absent; ``tools/utils.py`` imports it at module scope, so empty stub
     modules are injected (needed only for ``eval.matching``).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get('IMP_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'nets', 'gm.py'))


def load():
    """Returns a namespace with GM, DGNNS, AdaGMN, layers, matching (lazy), patched for CPU."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    # The reference's top-level package is called ``nets`` -- same name as our drop-in package.
    # Make sure the reference's wins for this import and is not shadowed by dropin/.
    for m in [k for k in sys.modules if k == 'nets' or k.startswith('nets.')]:
        del sys.modules[m]
    sys.path.insert(0, REF_ROOT)
    try:
        for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.cm'):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules['matplotlib'].use = lambda *a, **k: None
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
        sys.modules['matplotlib'].cm = sys.modules['matplotlib.cm']
        import nets.layers as ref_layers
        import nets.gm as ref_gm
        import nets.gms as ref_gms
        import nets.adgm as ref_adgm
    finally:
        sys.path.remove(REF_ROOT)

    from oracle import imp_oracle

    def sink_algorithm_any_device(M, dustbin, iteration):
        # dustbin padding + marginals rebuilt on M.device; the iteration itself is the
        # reference's own device-agnostic ``sinkhorn`` (nets/layers.py:27-35), unmodified.
        ma = imp_oracle.pad_dustbin(M, dustbin)
        r = ma.new_ones(ma.shape[0], ma.shape[1]); r[:, -1] = ma.shape[1]
        c = ma.new_ones(ma.shape[0], ma.shape[2]); c[:, -1] = ma.shape[2]
        return ref_layers.sinkhorn(ma, r, c, iteration)

    ref_gm.sink_algorithm = sink_algorithm_any_device
    ns = types.SimpleNamespace(GM=ref_gm.GM, DGNNS=ref_gms.DGNNS, AdaGMN=ref_adgm.AdaGMN,
                               layers=ref_layers, gm=ref_gm, root=REF_ROOT)
    return ns


def load_matching(ns):
    """eval.matching drivers (need cv2 + the matplotlib stub)."""
    sys.path.insert(0, REF_ROOT)
    try:
        import eval.matching as ref_matching
    finally:
        sys.path.remove(REF_ROOT)
    return ref_matching
