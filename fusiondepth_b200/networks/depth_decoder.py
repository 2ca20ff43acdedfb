"""U-Net disparity decoder on the package's NHWC kernels.

Mirrors ``networks.DepthDecoder`` (reference networks/depth_decoder.py:6-96): same constructor
flags (cat2end, road, catxy, deep), ``forward(input_features, two_channel, beam_features,
depth_maps, tanh)``, output dict keyed ("disp", s) and state-dict keys ``decoder.{k}...``.
The nearest-x2 upsample, the skip add (features + beam features), the channel concat and the
reflection pad are one gather kernel (fd_assemble_fwd); ELU / sigmoid / tanh run in the conv
epilogue.
"""
from __future__ import absolute_import, division, print_function

from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ..layers import ConvBlock, Conv3x3


class _DeepBlock(nn.Sequential):
    """Two ConvBlocks (`deep=True`, depth_decoder.py:27-33): keys ``{0,1}.conv.conv.*``."""

    def forward(self, x, segments=None):
        # the first block keeps its zero-padded output channels (262 -> 288 ...), the second consumes them
        return self[1](self[0](x, segments=segments, pad_out=True))


class DepthDecoder(nn.Module):
    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True,
                 cat2end=False, road=False, catxy=False, deep=False):
        super(DepthDecoder, self).__init__()
        self.num_output_channels = num_output_channels
        self.use_skips = use_skips
        self.upsample_mode = "nearest"
        self.scales = scales
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array([16, 32, 64, 128, 256])
        self.cat2end = cat2end

        def block(cin, cout):
            if deep:
                return _DeepBlock(ConvBlock(cin, cin), ConvBlock(cin, cout))
            return ConvBlock(cin, cout)

        self.convs = OrderedDict()
        for i in range(4, -1, -1):
            cin = self.num_ch_enc[-1] if i == 4 else self.num_ch_dec[i + 1]
            self.convs[("upconv", i, 0)] = block(cin, self.num_ch_dec[i])
            cin = self.num_ch_dec[i]
            if self.use_skips and i > 0:
                cin += self.num_ch_enc[i - 1]
            if road and i in self.scales and self.use_skips:
                cin += 6 if catxy else 3
            self.convs[("upconv", i, 1)] = block(cin, self.num_ch_dec[i])
        for s in self.scales:
            self.convs[("dispconv", s)] = Conv3x3(self.num_ch_dec[s], self.num_output_channels)
        if self.cat2end:
            self.convs[("dispconv", 0)] = Conv3x3(self.num_ch_dec[0] + 2, self.num_output_channels)
        self.decoder = nn.ModuleList(list(self.convs.values()))
        self.sigmoid = nn.Sigmoid()

    def forward(self, input_features, two_channel=None, beam_features=None, depth_maps=None,
                tanh=False):
        self.outputs = {}
        bf = beam_features
        # first ConvBlock consumes (f4 + beam f4) straight from the assemble gather
        segs = [(input_features[-1], bf[-1] if bf is not None else None, False)]
        x = None
        for i in range(4, -1, -1):
            x = self.convs[("upconv", i, 0)](x, segments=segs)
            segs = [(x, None, True)]
            if self.use_skips and i > 0:
                segs.append((input_features[i - 1], bf[i - 1] if bf is not None else None, False))
            if depth_maps is not None and i in self.scales and self.use_skips:
                segs.append((depth_maps[("disp", i)], None, False))
            x = self.convs[("upconv", i, 1)](None, segments=segs)
            segs = [(x, None, False)]
            if i in self.scales:
                if i == 0 and self.cat2end:
                    self.outputs[("disp", i)] = self.convs[("dispconv", i)](
                        None, act="sigmoid", segments=[(x, None, False), (two_channel, None, False)])
                else:
                    self.outputs[("disp", i)] = self.convs[("dispconv", i)](
                        x, act="tanh" if tanh else "sigmoid")
        return self.outputs
