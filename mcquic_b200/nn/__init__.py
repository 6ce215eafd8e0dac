from .blocks import *  # noqa: F401,F403
from .convs import *  # noqa: F401,F403
from .gdn import *  # noqa: F401,F403
