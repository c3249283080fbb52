"""shinnosuke_b200 — the data-parallel training hot path of E1eveNn/shinnosuke, rebuilt for
NVIDIA B200 (sm_100a): the reference's Keras-like Layer / Sequential / Model API over
hand-written CUDA kernels reached through a C ABI (include/shinnosuke_b200.h)."""
from .layers import *      # noqa: F401,F403
from .models import *      # noqa: F401,F403
from .models import Sequential, Model
from .utils import *       # noqa: F401,F403

name = 'shinnosuke-b200'
__version__ = '0.1.0'
