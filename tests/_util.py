"""Shared helpers for the test-suite (deterministic weights, tolerances, paths)."""
from __future__ import annotations

import hashlib
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _seed_of(key: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(("%d/%s" % (seed, key)).encode()).digest()[:4], "little")


def synth_weights(template: "OrderedDict[str, torch.Tensor]", seed: int = 0):
    """Deterministic weights for a state_dict *template* (only key -> shape/dtype is used),
    independent of module construction order: every tensor is drawn from its own generator
    seeded by hash(seed, key).  Conv/linear weights ~ N(0, 1/fan_in) (unit gain, so that the
    activations -- and the sigmoid logits -- stay O(1) and the nets are well conditioned in both
    train and eval mode); BN weight ~ U(.5,1.5);
    biases, running_mean ~ small normal; running_var ~ U(.5,1.5)."""
    out = OrderedDict()
    for k, t in template.items():
        g = torch.Generator().manual_seed(_seed_of(k, seed))
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(t.shape, generator=g)
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(t.shape, generator=g)
        elif t.dim() >= 2:
            fan_in = int(np.prod(t.shape[1:]))
            out[k] = torch.randn(t.shape, generator=g) * float(np.sqrt(1.0 / fan_in))
        elif ".bn" in k or "downsample.1" in k:
            out[k] = (0.5 + torch.rand(t.shape, generator=g)) if k.endswith("weight") \
                else 0.1 * torch.randn(t.shape, generator=g)
        else:
            out[k] = 0.05 * torch.randn(t.shape, generator=g)
    return out


def clone_sd(sd, requires_grad=False):
    out = OrderedDict()
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and not (
                k.endswith("running_mean") or k.endswith("running_var")):
            t.requires_grad_(True)
        out[k] = t
    return out


def rel_err(a, b) -> float:
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def coarse_disparity(B, H, W, seed=3):
    """A smooth synthetic stage-1 disparity [B,1,H,W] in (0,1) for the pseudo-3D pack fixtures."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 1, H // 8, W // 8, generator=g)
    d = torch.sigmoid(4 * (F.interpolate(low, (H, W), mode="bilinear", align_corners=False) - 0.5) - 1.0)
    return (d + 0.01 * torch.rand(B, 1, H, W, generator=g)).contiguous()
