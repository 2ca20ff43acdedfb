"""CPU restatement of the colour side of the reference's data producer.

TEST INFRASTRUCTURE ONLY -- nothing under fusiondepth_b200/ imports this; the product path is
fusiondepth_b200/dataprep.py on csrc/dataprep.cu.

The reference builds its colour inputs with PIL / torchvision on uint8 images (datasets/mono_dataset.py:85-104,
156-206): optional horizontal flip, a pyramid of `transforms.Resize(.., interpolation=Image.ANTIALIAS)` (each
scale resized from the previous one), `ColorJitter` on the PIL image, `ToTensor`.  The arithmetic underneath is
a third-party dependency that is not under /root/reference -- Pillow (12.2 in this image; Resample.c, Blend.c,
Convert.c) and torchvision 0.26 `transforms._functional_pil` -- restated here from their published algorithms:

  * resize: separable, horizontal pass first, uint8 intermediate; per output sample a window of
    `support = 3 * max(scale, 1)` input samples weighted by the Lanczos-3 kernel, weights normalised in double
    and rounded to 22-bit fixed point, accumulated in int32 from 1 << 21, shifted and clipped to uint8;
  * brightness / contrast / saturation: `Image.blend(degenerate, image, factor)` in float32 with truncation
    (clipping only outside [0, 1]); degenerate = black / the rounded mean of the L image / the L image, with
    L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16;
  * hue: RGB -> HSV (uint8, colorsys formulas in float32), h += uint8(factor * 255) with wrap-around, HSV -> RGB;
  * ToTensor: float32(u8) / 255.

Pinned by tests/test_oracle_data.py against PIL itself (random images and the reference's sizes, every
(value, factor) pair of the blends, all 2^24 colours through the hue round trip).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        if x == 0.0:
            return 1.0
        a = x * math.pi
        b = (x / 3.0) * math.pi
        return (math.sin(a) / a) * (math.sin(b) / b)
    return 0.0


def lanczos_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the whole-image box: (ksize, bounds [out,2] int32
    (first input sample, count), coefficients [out, ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    in_size = img.shape[axis]
    ksize, bounds, kk = lanczos_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_lanczos(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img [H,W,C] uint8 -> [out_h,out_w,C] uint8, PIL `Image.resize((out_w, out_h), Image.LANCZOS)`."""
    H, W = img.shape[:2]
    cur = img
    if out_w != W:
        cur = _resample_axis(cur, out_w, 1)
    if out_h != H:
        cur = _resample_axis(cur, out_h, 0)
    return cur


def to_L(img: np.ndarray) -> np.ndarray:
    r, g, b = (img[..., i].astype(np.uint32) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(deg: np.ndarray, img: np.ndarray, alpha: float) -> np.ndarray:
    """PIL Image.blend(degenerate, image, alpha) on uint8 arrays."""
    if alpha == 0.0:
        return deg.copy()
    if alpha == 1.0:
        return img.copy()
    a = np.float32(alpha)
    d = deg.astype(np.int32)
    t = (d.astype(np.float32) + a * (img.astype(np.int32) - d).astype(np.float32)).astype(np.float32)
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.int32).astype(np.uint8)          # C cast: truncation
    return np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32))).astype(np.uint8)


def adjust_brightness(img, f):
    return blend(np.zeros_like(img), img, f)


def adjust_contrast(img, f):
    L = to_L(img)
    mean = int(float(L.astype(np.int64).sum()) / L.size + 0.5)
    return blend(np.full_like(img, mean), img, f)


def adjust_saturation(img, f):
    L = to_L(img)
    return blend(np.repeat(L[..., None], 3, axis=-1), img, f)


def rgb_to_hsv(img: np.ndarray) -> np.ndarray:
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    f32 = np.float32
    cr = (maxc - minc).astype(f32)
    safe = np.where(cr == 0, f32(1), cr)
    s = cr / np.where(maxc == 0, 1, maxc).astype(f32)
    rc = (maxc - r).astype(f32) / safe
    gc = (maxc - g).astype(f32) / safe
    bc = (maxc - b).astype(f32) / safe
    # `2.0 + rc - bc` has double constants in C: evaluated in double, stored to float
    f64 = np.float64
    h = np.where(r == maxc, bc - gc,
                 np.where(g == maxc, (2.0 + rc.astype(f64) - bc.astype(f64)).astype(f32),
                          (4.0 + gc.astype(f64) - rc.astype(f64)).astype(f32))).astype(f32)
    # h = fmod(h / 6.0 + 1.0, 1.0) in double (the C constants are doubles), back to float
    hd = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(f32)
    uh = np.clip((hd.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    grey = minc == maxc
    out = np.stack([np.where(grey, 0, uh), np.where(grey, 0, us), maxc], axis=-1)
    return out.astype(np.uint8)


def hsv_to_rgb(hsv: np.ndarray) -> np.ndarray:
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    f32 = np.float32
    hf = h.astype(f32).astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int32)
    f = (hf - i.astype(f32).astype(np.float64)).astype(f32)
    fs = (s.astype(f32).astype(np.float64) / 255.0).astype(f32)
    vf = v.astype(f32)
    one = np.float64(1.0)
    p = np.round(vf.astype(np.float64) * (one - fs.astype(np.float64)))
    q = np.round(vf.astype(np.float64) * (one - fs.astype(np.float64) * f.astype(np.float64)))
    t = np.round(vf.astype(np.float64) * (one - fs.astype(np.float64) * (one - f.astype(np.float64))))
    p, q, t = (np.clip(x.astype(np.int32), 0, 255) for x in (p, q, t))
    sel = i % 6
    r = np.choose(sel, [v, q, p, p, t, v])
    g = np.choose(sel, [t, v, v, q, p, p])
    b = np.choose(sel, [p, p, t, v, v, q])
    grey = s == 0
    out = np.stack([np.where(grey, v, r), np.where(grey, v, g), np.where(grey, v, b)], axis=-1)
    return out.astype(np.uint8)


def adjust_hue(img, f):
    hsv = rgb_to_hsv(img)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + int(np.int32(f * 255).astype(np.uint8))).astype(np.uint8)
    return hsv_to_rgb(hsv)


OPS = {0: adjust_brightness, 1: adjust_contrast, 2: adjust_saturation, 3: adjust_hue}


def color_jitter(img, order, factors):
    """torchvision ColorJitter.forward with explicit parameters: `order` = permutation of (0 brightness,
    1 contrast, 2 saturation, 3 hue), factors[i] = factor of op i (None = skipped)."""
    for op in order:
        if factors[op] is not None:
            img = OPS[op](img, factors[op])
    return img


def to_tensor(img: np.ndarray, flip: bool = False) -> np.ndarray:
    """transforms.ToTensor: [H,W,3] uint8 -> [3,H,W] float32 in [0,1] (after the optional horizontal flip)."""
    if flip:
        img = img[:, ::-1]
    return (img.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1).copy()


def color_pyramid(native: np.ndarray, height: int, width: int, num_scales: int = 4, flip: bool = False,
                  jitter=None):
    """mono_dataset.preprocess for one frame: {("color", s)}, {("color_aug", s)} float tensors.
    jitter = None or a list (one per scale) of (order, factors)."""
    if flip:
        native = native[:, ::-1].copy()
    out = {}
    cur = native
    for s in range(num_scales):
        cur = resize_lanczos(cur, height >> s, width >> s)
        out[("color", s)] = to_tensor(cur)
        aug = cur if jitter is None else color_jitter(cur, *jitter[s])
        out[("color_aug", s)] = to_tensor(aug)
    return out
