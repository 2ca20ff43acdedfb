"""`import layers` shim: put fusiondepth_b200/dropin first on sys.path."""
from fusiondepth_b200.layers import *            # noqa: F401,F403
from fusiondepth_b200.layers import F, nn, np, torch  # noqa: F401  (part of the surface)
