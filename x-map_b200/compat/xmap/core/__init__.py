from xmap_b200.core import *  # noqa: F401,F403
