"""rollout-bayesian-optimization_b200: B200-native (sm_100a, FP64 CUDA) Monte-Carlo rollout acquisition estimator
and adjoint gradient, behind the reference's own interface names (see api.py) and a C ABI (include/rbo.h)."""
from . import _lib, api, problems  # noqa: F401
from .api import *  # noqa: F401,F403
from ._lib import Handle, RboError, LIB_PATH  # noqa: F401
