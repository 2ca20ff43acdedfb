"""fusiondepth_b200: B200-native (sm_100a) implementation of FusionDepth's per-step training
hot path behind the reference's ``layers`` / ``networks`` module surface.

    from fusiondepth_b200 import layers, networks, lidar, training

``fusiondepth_b200/dropin`` can be put on ``sys.path`` so that the reference's unchanged
``trainer.py`` resolves ``import networks`` / ``from layers import *`` to this package.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (ctypes binding; loading is deferred to first use)
