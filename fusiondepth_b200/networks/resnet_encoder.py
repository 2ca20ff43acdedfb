"""ResNet-18/34/50/101/152 trunk on the package's NHWC kernels.

Mirrors ``networks.ResnetEncoder`` (reference networks/resnet_encoder.py:53-103): same
constructor, ``num_ch_enc``, ``.encoder`` attribute and state-dict keys
(``encoder.conv1.weight``, ``encoder.bn1.*``, ``encoder.layer{1-4}.{i}.{conv,bn}{1,2,3}.*``,
``encoder.layer*.0.downsample.{0,1}.*``, ``encoder.fc.*`` -- fc/avgpool are kept although unused,
SURVEY.md Appendix E.9) so checkpoints interchange with the reference.  torchvision is not
needed: block wiring follows torchvision's BasicBlock / Bottleneck (v1.5, stride on the 3x3).
"""
from __future__ import absolute_import, division, print_function

import numpy as np
import torch
import torch.nn as nn

from .. import ops

BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3),
          152: (3, 8, 36, 3)}


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameters/initialisation, forward on fd_conv2d_* (weights kept channels-last)."""

    def __init__(self, *a, **k):
        super(Conv2d, self).__init__(*a, **k)
        self.weight.data = self.weight.data.contiguous(memory_format=torch.channels_last)

    def forward(self, x, act="none", stats=None):
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0], act, stats)


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d state, forward on fd_bn_* with optional fused residual add and ReLU."""

    def forward(self, x, residual=None, relu=False, stats=None):
        training = self.training or not self.track_running_stats
        stat_weight = -1.0
        if self.training and self.track_running_stats:
            sched = ops.BN_SCHEDULE
            if sched is not None and self.momentum is not None:
                stat_weight = sched.next_weight(self)      # the step bumps num_batches_tracked itself
            if stat_weight < 0:
                self.num_batches_tracked += 1
        return ops.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var,
                              residual, training, self.momentum, self.eps, relu, stat_weight, stats)


def conv_bn(conv, bn, x, residual=None, relu=False):
    """bn(conv(x)) (+ residual, ReLU).  With ops.FUSE_BN_STATS the batch statistics come out of the
    convolution's epilogue (a zeroed fp64 [2*C] buffer travels from the conv to the BatchNorm)."""
    ws = None
    if (bn.training or not bn.track_running_stats) and ops.conv_emits_stats(
            conv.in_channels, conv.out_channels, conv.bias is not None):
        ws = ops.zeroed_stats(2 * conv.out_channels, x.device)
    return bn(conv(x, stats=ws), residual=residual, relu=relu, stats=ws)


class _Downsample(nn.Sequential):
    def forward(self, x):
        return conv_bn(self[0], self[1], x)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        self.conv1 = Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = conv_bn(self.conv1, self.bn1, x, relu=True)
        identity = x if self.downsample is None else self.downsample(x)
        return conv_bn(self.conv2, self.bn2, out, residual=identity, relu=True)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super(Bottleneck, self).__init__()
        self.conv1 = Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = conv_bn(self.conv1, self.bn1, x, relu=True)
        out = conv_bn(self.conv2, self.bn2, out, relu=True)
        identity = x if self.downsample is None else self.downsample(x)
        return conv_bn(self.conv3, self.bn3, out, residual=identity, relu=True)


class ResNetTrunk(nn.Module):
    """conv1/bn1/maxpool/layer1-4 (+ unused avgpool, fc) with torchvision's attribute names."""

    def __init__(self, num_layers, in_channels=3):
        super(ResNetTrunk, self).__init__()
        block = BasicBlock if num_layers < 50 else Bottleneck
        self.inplanes = 64
        self.conv1 = Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, BLOCKS[num_layers][0])
        self.layer2 = self._make_layer(block, 128, BLOCKS[num_layers][1], stride=2)
        self.layer3 = self._make_layer(block, 256, BLOCKS[num_layers][2], stride=2)
        self.layer4 = self._make_layer(block, 512, BLOCKS[num_layers][3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, 1000)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = _Downsample(
                Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)


def find_imagenet_checkpoint(num_layers):
    """Path of torchvision's ``resnet{N}-<hash>.pth`` on local disk.  The reference downloads it
    (model_zoo.load_url, resnet_encoder.py:46 / torchvision's ``pretrained=True``); this package never opens a
    socket, so the file has to be in $FD_PRETRAINED_DIR or in torch.hub's checkpoint cache already."""
    import glob
    import os
    dirs = [os.environ.get("FD_PRETRAINED_DIR"),
            os.path.join(torch.hub.get_dir(), "checkpoints")]
    for d in dirs:
        if d:
            hits = sorted(glob.glob(os.path.join(d, "resnet%d-*.pth" % num_layers)))
            if hits:
                return hits[0]
    raise RuntimeError(
        "pretrained=True: no resnet%d-*.pth under %s; copy torchvision's ImageNet checkpoint there "
        "(or set FD_PRETRAINED_DIR), or construct with pretrained=False (--weights_init scratch) and "
        "load a checkpoint via load_state_dict" % (num_layers, [d for d in dirs if d]))


def load_imagenet_weights(trunk, num_layers, num_input_images=1, keep_conv1=True):
    """Fill a ResNetTrunk from torchvision's ImageNet state dict (same keys).  conv1 is repeated over the
    stacked input frames and divided by their number (resnet_encoder.py:47-48); with keep_conv1=False the
    trunk's own (replaced, randomly initialised) conv1 stays."""
    loaded = torch.load(find_imagenet_checkpoint(num_layers), map_location="cpu", weights_only=True)
    loaded = dict(loaded)
    if keep_conv1:
        if num_input_images > 1:
            loaded["conv1.weight"] = torch.cat([loaded["conv1.weight"]] * num_input_images, 1) / num_input_images
    else:
        loaded["conv1.weight"] = trunk.conv1.weight.detach()
    trunk.load_state_dict(loaded)
    for m in trunk.modules():        # load_state_dict copies into the existing storage; keep the NHWC layout explicit
        if isinstance(m, nn.Conv2d):
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)


class ResnetEncoder(nn.Module):
    """Pytorch-module-compatible ResNet encoder (reference resnet_encoder.py:53-103)."""

    def __init__(self, num_layers, pretrained, num_input_images=1, cat4beam_to_color=False,
                 cat2channel=False, beam_encoder=False, refine_encoder=False):
        super(ResnetEncoder, self).__init__()
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        if num_layers not in BLOCKS:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        # conv1 input channels: resnet_encoder.py:71-87
        cin = 3 * num_input_images if num_input_images > 1 else 3
        if cat4beam_to_color:
            cin = 4
        elif cat2channel:
            cin = 5
        elif beam_encoder:
            cin = 2 * num_input_images if num_input_images > 1 else 2
        elif refine_encoder:
            cin = 6
        self.num_layers = num_layers
        self.encoder = ResNetTrunk(num_layers, cin)
        if pretrained:
            # resnet_encoder.py:45-49, 71-87: the ImageNet weights go in first (conv1 tiled over the stacked
            # frames), then the 4/5/2/6-channel variants replace conv1 by a freshly initialised one
            keep_conv1 = cin == (3 * num_input_images if num_input_images > 1 else 3)
            load_imagenet_weights(self.encoder, num_layers, num_input_images, keep_conv1)
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4

    def forward(self, input_image):
        enc = self.encoder
        self.features = []
        # (x - 0.45) / 0.225 and the 7x7/2 stem in one op (normalised im2col rows -> GEMM)
        self.features.append(enc.bn1(ops.stem_conv(input_image, enc.conv1.weight), relu=True))
        x = ops.maxpool3x3s2(self.features[-1])
        for li, layer in enumerate((enc.layer1, enc.layer2, enc.layer3, enc.layer4)):
            if li >= 2:
                # bucketed gradient exchange: layer4's (then layer3's) gradients are complete once the
                # backward pass reaches this point (training.bucket_of: layer4 -> 0, layer3 -> 1)
                x = ops.grad_ready(x, 3 - li)
            for blk in layer:
                x = blk(x)
            self.features.append(x)
        return self.features
