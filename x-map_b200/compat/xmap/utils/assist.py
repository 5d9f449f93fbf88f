from xmap_b200.utils.assist import *  # noqa: F401,F403
