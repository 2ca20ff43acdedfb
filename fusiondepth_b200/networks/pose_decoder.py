"""Pose head on the package's kernels.

Mirrors ``networks.PoseDecoder`` (reference networks/pose_decoder.py:8-51): squeeze 1x1 -> two
3x3 -> 1x1 -> spatial mean -> x0.01 -> (axisangle, translation); state-dict keys ``net.{0..3}.*``.
ReLU runs in the conv epilogue; the mean and the 0.01 scale are one kernel.
"""
from __future__ import absolute_import, division, print_function

from collections import OrderedDict

import torch
import torch.nn as nn

from .. import ops
from .resnet_encoder import Conv2d


class PoseDecoder(nn.Module):
    def __init__(self, num_ch_enc, num_input_features, num_frames_to_predict_for=None, stride=1):
        super(PoseDecoder, self).__init__()
        self.num_ch_enc = num_ch_enc
        self.num_input_features = num_input_features
        if num_frames_to_predict_for is None:
            num_frames_to_predict_for = num_input_features - 1
        self.num_frames_to_predict_for = num_frames_to_predict_for
        self.convs = OrderedDict()
        self.convs[("squeeze")] = Conv2d(int(self.num_ch_enc[-1]), 256, 1)
        self.convs[("pose", 0)] = Conv2d(num_input_features * 256, 256, 3, stride, 1)
        self.convs[("pose", 1)] = Conv2d(256, 256, 3, stride, 1)
        self.convs[("pose", 2)] = Conv2d(256, 6 * num_frames_to_predict_for, 1)
        self.relu = nn.ReLU()
        self.net = nn.ModuleList(list(self.convs.values()))

    def forward(self, input_features, beam_inputs=None):
        if beam_inputs is not None:
            last_features = [ops.add(input_features[0][-1], beam_inputs[0][-1])]
        else:
            last_features = [f[-1] for f in input_features]
        cat = [self.convs["squeeze"](f, act="relu") for f in last_features]
        out = cat[0] if len(cat) == 1 else ops.assemble([(c, None, False) for c in cat], pad=0)
        for i in range(3):
            out = self.convs[("pose", i)](out, act="relu" if i != 2 else "none")
        out = ops.mean_hw(out, 0.01).view(-1, self.num_frames_to_predict_for, 1, 6)
        return out[..., :3], out[..., 3:]
