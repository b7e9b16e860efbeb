"""Drop-in `nets` package: put this directory in front of the reference tree on sys.path and the reference's
eval/eval_imp.py / eval/matching.py import the B200 implementation instead of its own nets/*.py
(INTEGRATION.md).  Pure re-exports; the implementation lives in imp_release_b200."""
import os
import sys

_repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _repo not in sys.path:
    sys.path.insert(0, _repo)

# submodules this package does not provide (nets.superglue, nets.loss, ...) resolve to the reference tree's own nets/
_here = os.path.dirname(os.path.abspath(__file__))
__path__ = [_here] + [os.path.join(p or '.', 'nets') for p in sys.path
                      if os.path.isdir(os.path.join(p or '.', 'nets')) and os.path.abspath(os.path.join(p or '.', 'nets')) != _here]
