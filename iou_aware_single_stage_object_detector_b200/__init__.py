"""B200-native IoU-aware RetinaNet inference path (sm_100a CUDA behind a C ABI)."""
from . import lib
from .api import *  # noqa: F401,F403
from .api import Config, build_detector  # noqa: F401

__version__ = "0.1.0"
