"""oracle/data_oracle.py (the restatement of Pillow's resize / blend / HSV arithmetic behind the reference's data
producer, mono_dataset.py:85-104) pinned against PIL and torchvision themselves."""
import numpy as np
import pytest

from oracle import data_oracle as D

Image = pytest.importorskip("PIL.Image")
FP = pytest.importorskip("torchvision.transforms._functional_pil")


def _img(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    base[: h // 3] = (np.linspace(0, 255, w)[None, :, None] + rng.integers(0, 20, (h // 3, w, 3))).clip(0, 255)
    return base


@pytest.mark.parametrize("shape", [(375, 1242, 192, 640), (192, 640, 96, 320), (96, 320, 48, 160), (48, 160, 24, 80),
                                   (370, 1226, 320, 1024), (100, 130, 37, 51), (64, 64, 64, 32), (33, 47, 66, 94)])
def test_resize_matches_pil(shape):
    H, W, h, w = shape
    img = _img(H, W, H + w)
    ref = np.array(Image.fromarray(img).resize((w, h), Image.LANCZOS))
    assert np.array_equal(D.resize_lanczos(img, h, w), ref)


def test_blends_match_pil():
    rng = np.random.default_rng(3)
    img = np.zeros((256, 256, 3), np.uint8)
    img[..., 0] = np.arange(256)[:, None]
    img[..., 1] = np.arange(256)[None, :]
    img[..., 2] = rng.integers(0, 256, (256, 256))
    for f in [float(v) for v in np.linspace(0.8, 1.2, 21)] + [0.0, 1.0, 0.5, 1.7, 1 / 3]:
        for fn, pf in ((D.adjust_brightness, FP.adjust_brightness), (D.adjust_contrast, FP.adjust_contrast),
                       (D.adjust_saturation, FP.adjust_saturation)):
            assert np.array_equal(fn(img, f), np.array(pf(Image.fromarray(img), f))), (fn.__name__, f)


def test_hue_matches_pil_on_a_colour_lattice():
    # every colour with channel values on a 5-step lattice through 255 plus 2^20 random ones (all 2^24 pass too: ~40 s)
    v = np.unique(np.concatenate([np.arange(0, 256, 5), [1, 2, 127, 128, 254, 255]])).astype(np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    lattice = np.stack([r, g, b], -1).reshape(-1, 1, 3)
    rnd = np.random.default_rng(5).integers(0, 256, (1 << 20, 1, 3), dtype=np.uint8)
    img = np.concatenate([lattice, rnd])
    img = img[: img.shape[0] // 64 * 64].reshape(-1, 64, 3)
    assert np.array_equal(D.rgb_to_hsv(img), np.array(Image.fromarray(img).convert("HSV")))
    assert np.array_equal(D.hsv_to_rgb(img), np.array(Image.fromarray(img, "HSV").convert("RGB")))
    for f in (-0.1, -0.037, 0.05, 0.1, 0.5):
        assert np.array_equal(D.adjust_hue(img, f), np.array(FP.adjust_hue(Image.fromarray(img), f))), f


def test_pyramid_matches_the_reference_preprocess_recipe():
    import torchvision.transforms as T
    native = _img(375, 1242, 9)
    pil = Image.fromarray(native).transpose(Image.FLIP_LEFT_RIGHT)       # get_color(..., do_flip=True)
    want = {}
    cur = pil
    for s in range(4):
        cur = T.Resize((192 >> s, 640 >> s), interpolation=T.InterpolationMode.LANCZOS)(cur)   # mono_dataset.py:76-79
        want[s] = T.ToTensor()(cur).numpy()
    got = D.color_pyramid(native, 192, 640, 4, flip=True)
    for s in range(4):
        assert np.array_equal(got[("color", s)], want[s]), s
