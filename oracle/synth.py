"""`oracle.synth` = the repo-level synthetic data generators (synth_data.py), re-exported so that the oracle, the golden
generator and the tests keep one import path.  bench.py's GPU arm and tools/ import `synth_data` directly: they use
nothing of the oracle."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from synth_data import *  # noqa: F401,F403,E402
from synth_data import _gen, _trunc_normal, _block_weights  # noqa: F401,E402
