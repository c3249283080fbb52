from .Activator import *        # noqa: F401,F403
from .Initializers import *     # noqa: F401,F403
from .MiniBatch import *        # noqa: F401,F403
from .Preprocess import *       # noqa: F401,F403
from .Optimizers import *       # noqa: F401,F403
from .Objectives import *       # noqa: F401,F403
from .Regularizers import *     # noqa: F401,F403

name = 'shinnosuke-b200-utils'
