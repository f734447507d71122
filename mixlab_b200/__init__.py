"""mixlab_b200 -- B200 (sm_100a) back end for Mixlab's per-tick module-graph hot path.

The product is the C-ABI shared library libmixlab_b200.so (include/mixlab_b200.h) built from
mixlab_b200/csrc; `mixlab_b200.api` is its ctypes view.  Nothing here falls back to a CPU path.
"""
from . import api  # noqa: F401
from .api import *  # noqa: F401,F403
