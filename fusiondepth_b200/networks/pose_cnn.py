"""Seven-conv pose network (``--pose_model_type posecnn``) on the package's kernels.

Mirrors ``networks.PoseCNN`` (reference networks/pose_cnn.py:7-44); state-dict keys
``net.{0..6}.*`` and ``pose_conv.*``.
"""
from __future__ import absolute_import, division, print_function

import torch
import torch.nn as nn

from .. import ops
from .resnet_encoder import Conv2d


class PoseCNN(nn.Module):
    def __init__(self, num_input_frames):
        super(PoseCNN, self).__init__()
        self.num_input_frames = num_input_frames
        spec = [(3 * num_input_frames, 16, 7), (16, 32, 5), (32, 64, 3), (64, 128, 3),
                (128, 256, 3), (256, 256, 3), (256, 256, 3)]
        self.convs = {i: Conv2d(ci, co, k, 2, (k - 1) // 2) for i, (ci, co, k) in enumerate(spec)}
        self.pose_conv = Conv2d(256, 6 * (num_input_frames - 1), 1)
        self.num_convs = len(self.convs)
        self.relu = nn.ReLU(True)
        self.net = nn.ModuleList(list(self.convs.values()))

    def forward(self, out):
        # raw NCHW images in; no input normalisation here (pose_cnn.py:31-35)
        out = out.contiguous(memory_format=torch.channels_last)
        for i in range(self.num_convs):
            out = self.convs[i](out, act="relu")
        out = self.pose_conv(out)
        out = ops.mean_hw(out, 0.01).view(-1, self.num_input_frames - 1, 1, 6)
        return out[..., :3], out[..., 3:]
