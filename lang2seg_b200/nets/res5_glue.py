"""res5 (`resnet.layer4`) as GLUE for the chained "second number" of the benchmark (SURVEY.md section 8 row a8,
8d "Step boundary").

Reference: `_head_to_tail` = `self.resnet.layer4(pool5)` (pyutils/mask-faster-rcnn/lib/nets/resnet_v1_cycle_response.py:271-273)
with layer4 = 3 Bottlenecks 1024 -> 2048, stride 1 (:131), 1x1 (stride) / 3x3 / 1x1 convolutions + BatchNorm each
(:79-113), downsample 1x1 + BatchNorm on the first block (:143-151).  The BatchNorm layers of the reference are frozen
(eval mode, no gradient: :343-357), i.e. per-channel affine maps -- they are folded into the convolutions here, which is
the same function and keeps 1/3 fewer activations alive.

This is NOT one of the path's kernels and is not re-implemented: dense convolutions on cuDNN, exactly as the reference
runs them.  It exists so that bench.py can report the reference-faithful chain
(crop -> res5 -> box head + mask head ; gated / ungated map -> res5 -> caption features -> att2in2) next to the graded
step that feeds each component synthetic inputs of the same shapes.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class FoldedBottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=False):
        super().__init__()
        out = planes * self.expansion
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=True)   # conv + frozen BN folded
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=True)
        self.conv3 = nn.Conv2d(planes, out, kernel_size=1, bias=True)
        self.downsample = nn.Conv2d(inplanes, out, kernel_size=1, stride=stride, bias=True) if downsample else None

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.conv1(x), inplace=True)
        y = F.relu(self.conv2(y), inplace=True)
        y = self.conv3(y)
        return F.relu(y + residual, inplace=True)


class Res5Glue(nn.Module):
    """layer4 of ResNet-101 with stride 1: (N,1024,h,w) -> (N,2048,h,w).  `chunk` bounds the rows per cuDNN call (and
    with it the workspace); activations of all chunks stay alive for the backward like in the reference."""

    def __init__(self, inplanes=1024, planes=512, blocks=3, chunk=4096):
        super().__init__()
        layers = [FoldedBottleneck(inplanes, planes, 1, downsample=True)]
        layers += [FoldedBottleneck(planes * 4, planes) for _ in range(blocks - 1)]
        self.layer4 = nn.Sequential(*layers)
        self.chunk = chunk
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                m.weight.data.mul_(0.5)          # keeps the 3-block stack at O(1) activations with folded BN
                nn.init.zeros_(m.bias)

    def forward(self, x):
        if x.shape[0] <= self.chunk:
            return self.layer4(x)
        return torch.cat([self.layer4(c) for c in x.split(self.chunk)], 0)
