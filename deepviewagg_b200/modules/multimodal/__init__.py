from .pooling import *  # noqa: F401,F403
from .fusion import *  # noqa: F401,F403
