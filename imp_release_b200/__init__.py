"""imp_release_b200 -- B200-native (sm_100a) implementation of the IMP / EIMP matching hot path.

Public API mirrors the reference's model layer (nets/gm.py::GM, nets/gms.py::DGNNS, nets/adgm.py::AdaGMN,
nets/layers.py::normalize_keypoints); ``dropin/nets`` re-exports it under the reference's module paths so that
``eval/eval_imp.py`` and ``eval/matching.py`` run unchanged.  All arithmetic is in ``libimp_b200.so``.
"""
from .nets.gm import GM, normalize_keypoints  # noqa: F401
from .nets.gms import DGNNS  # noqa: F401
from .nets.adgm import AdaGMN  # noqa: F401

__all__ = ['GM', 'DGNNS', 'AdaGMN', 'normalize_keypoints']
