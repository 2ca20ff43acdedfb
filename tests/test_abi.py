"""The C-ABI library loads and exports every symbol include/fusiondepth_b200.h declares."""
import ctypes
import os
import re

from tests._util import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fusiondepth_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("fd_photoloss_fwd", "fd_photoloss_bwd", "fd_lidar_depth_map", "fd_two_channel",
                 "fd_conv2d_fwd", "fd_conv2d_dgrad", "fd_conv2d_wgrad", "fd_bn_fwd", "fd_adam_step"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from fusiondepth_b200 import _lib, build
    so = build.build()
    lib = ctypes.CDLL(so)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export " + s
    # the python binding types exactly the declared set
    assert sorted(_lib.EXPORTS) == declared_symbols()
    lib2 = _lib.load()
    assert lib2.fd_version() >= 100
    assert lib2.fd_photoloss_workspace_bytes(2, 64, 96) > 0


def test_sass_is_sm100a():
    import subprocess
    from fusiondepth_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from fusiondepth_b200 import _lib, ops
    with pytest.raises(_lib.FusionDepthLibraryError):
        ops.conv2d(torch.zeros(1, 16, 8, 8), torch.zeros(16, 16, 3, 3))
