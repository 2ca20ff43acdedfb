"""CUDA LiDAR kernels vs the C oracle and the reference fixtures: bit-exact."""
import numpy as np
import pytest
import torch

from tests._util import GOLDEN
from fusiondepth_b200 import synth
from oracle import lidar_oracle as LO

pytestmark = pytest.mark.gpu


def _P(cam=2):
    return synth.velo_to_image_matrix(synth.parse_roundtrip(), cam)


@pytest.mark.parametrize("case", ["ring4", "piled", "dense", "edge"])
def test_depth_map_fixture_bit_exact(cuda, case):
    from fusiondepth_b200 import lidar
    g = np.load(GOLDEN + "/lidar.npz")
    pts = g[case + "/points"]
    d = lidar.depth_maps([pts], [_P()], 1242, 375, False, (384, 1280))
    assert np.array_equal(d[0].cpu().numpy(), g[case + "/depth384"])
    raw = lidar.depth_maps([pts], [_P()], 1242, 375, False, None)
    assert np.array_equal(raw[0].cpu().numpy(), g[case + "/depth_raw"])
    if case + "/depth_vel375" in g:
        dv = lidar.depth_maps([pts], [_P(3)], 1242, 375, True, (375, 1242))
        assert np.array_equal(dv[0].cpu().numpy(), g[case + "/depth_vel375"])
    fb = lidar.four_beam(d)
    assert np.array_equal(fb[0].cpu().numpy(), g[case + "/4beam"])
    tc = lidar.two_channel(fb)
    assert np.array_equal(tc[0].cpu().numpy(), g[case + "/2channel"])


@pytest.mark.parametrize("case", ["rand002", "rand015", "rand050"])
def test_two_channel_fixture_bit_exact(cuda, case):
    from fusiondepth_b200 import lidar
    g = np.load(GOLDEN + "/lidar.npz")
    tc = lidar.two_channel(torch.from_numpy(g[case + "/4beam"]).cuda()[None])
    assert np.array_equal(tc[0].cpu().numpy(), g[case + "/2channel"])


def test_batched_frames_vs_oracle(cuda):
    """many ragged frames in one launch, incl. an empty one and a 120k-point dense scan"""
    from fusiondepth_b200 import lidar
    scans = [synth.make_scan(100 + i, piled=50 * i) for i in range(5)]
    scans.append(np.zeros((0, 4), np.float32))
    scans.append(synth.make_dense_scan(42, n=120000))
    scans.append(synth.make_dense_scan(43, n=60000, edge_heavy=True))
    Ps = [_P(2 + (i % 2)) for i in range(len(scans))]
    d = lidar.depth_maps(scans, Ps, 1242, 375, False, (384, 1280))
    fb = lidar.four_beam(d)
    tc = lidar.two_channel(fb)
    for i, (p, P) in enumerate(zip(scans, Ps)):
        ref = LO.depth_map(p, P, 1242, 375, shape=(384, 1280))
        assert np.array_equal(d[i].cpu().numpy(), ref), i
        rfb = LO.pool_scale(ref)
        assert np.array_equal(fb[i].cpu().numpy(), rfb), i
        assert np.array_equal(tc[i].cpu().numpy(), LO.two_channel(rfb)), i


def test_dropin_signatures(cuda, tmp_path):
    from fusiondepth_b200 import lidar
    synth.write_calib_files(str(tmp_path))
    pts = synth.make_scan(9)
    fn = str(tmp_path / "0000000000.bin")
    pts.tofile(fn)
    out = lidar.generate_depth_map(str(tmp_path), fn, 2, shape=[384, 1280])
    assert out.dtype == np.float64
    ref = LO.depth_map(pts, _P(), 1242, 375, shape=(384, 1280))
    assert np.array_equal(out, ref)
    e, c = lidar.get_4beam_2channel(torch.from_numpy(LO.pool_scale(ref)))
    tc = LO.two_channel(LO.pool_scale(ref))
    assert np.array_equal(e.numpy(), tc[0]) and np.array_equal(c.numpy(), tc[1])


def test_other_resolution_window(cuda):
    """config-3 style 320x1024 maps (window scaled; reference data layer is 192x640-only)"""
    from fusiondepth_b200 import lidar
    g = torch.Generator().manual_seed(3)
    fb = ((torch.rand(2, 320, 1024, generator=g) < 0.03).float() * torch.rand(2, 320, 1024, generator=g))
    win = synth.lidar_window(320, 1024)
    tc = lidar.two_channel(fb.cuda(), win)
    for i in range(2):
        assert np.array_equal(tc[i].cpu().numpy(), LO.two_channel(fb[i].numpy(), win))
