from xmap_b200.core.baselinerSim import *  # noqa: F401,F403
from xmap_b200.core import baselinerSim as _m
globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith('__')})
