"""Drop-in `components` package: `components.readers` and `components.extractors` resolve to the B200 implementations, every
other submodule (evaluators, utils, ...) to the reference tree's own `components/` found further down sys.path."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(os.path.dirname(_here))
if _repo not in sys.path:
    sys.path.insert(0, _repo)
__path__ = [_here] + [os.path.join(p or '.', 'components') for p in sys.path
                      if os.path.isdir(os.path.join(p or '.', 'components'))
                      and os.path.abspath(os.path.join(p or '.', 'components')) != _here]
