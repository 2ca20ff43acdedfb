"""`import networks` shim: put fusiondepth_b200/dropin first on sys.path."""
from fusiondepth_b200.networks import ResnetEncoder, DepthDecoder, PoseDecoder, PoseCNN  # noqa: F401
