"""On-device data producer (csrc/dataprep.cu) against the CPU oracle and against PIL /
torchvision directly: bit-exact (uint8 / float32 array_equal)."""
import numpy as np
import pytest
import torch

from oracle import data_oracle as D

pytestmark = pytest.mark.gpu


def _img(B, h, w, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    x[:, : h // 3] = (np.linspace(0, 255, w)[None, None, :, None] + rng.integers(0, 20, (B, h // 3, w, 3))).clip(0, 255)
    return x


@pytest.mark.parametrize("shape", [(375, 1242, 192, 640), (192, 640, 96, 320), (48, 160, 24, 80),
                                   (370, 1226, 320, 1024), (100, 130, 37, 51), (33, 47, 66, 94), (64, 64, 64, 32)])
def test_resize_bit_exact(cuda, shape):
    from fusiondepth_b200 import dataprep
    H, W, h, w = shape
    x = _img(3, H, W, h)
    flip = np.array([0, 1, 0], np.uint8)
    got = dataprep.resize_lanczos(torch.from_numpy(x).cuda(), h, w, torch.from_numpy(flip)).cpu().numpy()
    for b in range(3):
        src = x[b, :, ::-1].copy() if flip[b] else x[b]
        assert np.array_equal(got[b], D.resize_lanczos(src, h, w)), b


def test_color_jitter_bit_exact(cuda):
    from fusiondepth_b200 import dataprep
    B, H, W = 12, 96, 320
    x = _img(B, H, W, 7)
    g = torch.Generator().manual_seed(11)
    order = torch.stack([torch.randperm(4, generator=g) for _ in range(B)]).to(torch.int32)
    lo = torch.tensor([0.8, 0.8, 0.8, -0.1], dtype=torch.float64)
    hi = torch.tensor([1.2, 1.2, 1.2, 0.1], dtype=torch.float64)
    f = lo + torch.rand(B, 4, generator=g, dtype=torch.float64) * (hi - lo)
    f[1, 0] = float("nan")                      # op skipped
    f[2, :3] = torch.tensor([0.0, 1.0, 1.7])    # blend corner cases: degenerate copy, image copy, clipping
    f[3, 3] = 0.5
    f[4, 3] = -0.5
    got = dataprep.color_jitter(torch.from_numpy(x).cuda(), order, f).cpu().numpy()
    for b in range(B):
        fac = [None if np.isnan(v) else float(np.float32(v)) if i < 3 else float(v) for i, v in enumerate(f[b].tolist())]
        want = D.color_jitter(x[b], order[b].tolist(), fac)
        assert np.array_equal(got[b], want), (b, order[b].tolist(), fac)


def test_hue_all_paths_bit_exact(cuda):
    from fusiondepth_b200 import dataprep
    v = np.unique(np.concatenate([np.arange(0, 256, 5), [1, 2, 127, 128, 254, 255]])).astype(np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    img = np.stack([r, g, b], -1).reshape(-1, 3)
    img = img[: img.shape[0] // 256 * 256].reshape(1, -1, 256, 3)
    order = torch.tensor([[3, 0, 1, 2]], dtype=torch.int32)
    for hue in (-0.1, 0.037, 0.1):
        f = torch.tensor([[float("nan"), float("nan"), float("nan"), hue]], dtype=torch.float64)
        got = dataprep.color_jitter(torch.from_numpy(img).cuda(), order, f).cpu().numpy()
        assert np.array_equal(got[0], D.adjust_hue(img[0], hue)), hue


def test_pyramid_matches_pil_and_torchvision(cuda):
    """The whole preprocess recipe of mono_dataset.py:85-104 for a flipped and an unflipped frame, against PIL /
    torchvision run here on the host (not against our own oracle)."""
    Image = pytest.importorskip("PIL.Image")
    import torchvision.transforms as T
    import torchvision.transforms._functional_pil as FP
    from fusiondepth_b200 import dataprep
    native = _img(2, 375, 1242, 21)
    pyr = dataprep.ColorPyramid(192, 640)
    order, factors = pyr.sample_jitter(2, torch.Generator().manual_seed(5))
    out = pyr(torch.from_numpy(native).cuda(), flip=torch.tensor([False, True]), jitter=(order, factors))
    ops = (FP.adjust_brightness, FP.adjust_contrast, FP.adjust_saturation, FP.adjust_hue)
    for b in range(2):
        cur = Image.fromarray(native[b])
        if b == 1:
            cur = cur.transpose(Image.FLIP_LEFT_RIGHT)
        for s in range(4):
            cur = T.Resize((192 >> s, 640 >> s), interpolation=T.InterpolationMode.LANCZOS)(cur)
            assert np.array_equal(out[("color", s)][b].cpu().numpy(), T.ToTensor()(cur).numpy()), (b, s)
            aug = cur
            for op in order[s, b].tolist():
                fac = float(factors[s, b, op])
                aug = ops[op](aug, float(np.float32(fac)) if op < 3 else fac)
            assert np.array_equal(out[("color_aug", s)][b].cpu().numpy(), T.ToTensor()(aug).numpy()), (b, s)
