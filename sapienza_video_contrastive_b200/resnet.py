"""The encoder stays on stock PyTorch / cuDNN (BASELINE.json north_star): torchvision ResNet-18/50 turned into the
stride-8 feature extractor the reference trains (code/resnet.py:17-54: layer3/layer4 stride 1, optional reflect
padding, fc / avgpool dropped) and wrapped so a (N,C,T,h,w) clip runs frame by frame (code/utils/__init__.py:285-297).
Module and parameter names match the reference so its checkpoints load unchanged (`encoder.model.<resnet keys>`)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torchvision.models.resnet as tv


class StrideEightResNet(tv.ResNet):
    def modify(self, remove_layers=(), padding=""):
        for name in ("layer3", "layer4"):
            layer = getattr(self, name, None)
            if layer is None:
                continue
            for m in layer.modules():
                if isinstance(m, nn.Conv2d):
                    m.stride = (1,) * len(m.stride)
        if padding:
            for m in self.modules():
                if isinstance(m, nn.Conv2d) and sum(m.padding) > 0:
                    m.padding_mode = padding
        for name in list(remove_layers) + ["fc", "avgpool"]:
            if getattr(self, name, None) is not None:
                setattr(self, name, None)

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        if self.maxpool is not None:
            x = self.maxpool(x)
        x = self.layer2(self.layer1(x))
        for name in ("layer3", "layer4"):
            layer = getattr(self, name)
            if layer is not None:
                x = layer(x)
        return x


def resnet18(**kw):
    return StrideEightResNet(tv.BasicBlock, [2, 2, 2, 2], **kw)


def resnet50(**kw):
    return StrideEightResNet(tv.Bottleneck, [3, 4, 6, 3], **kw)


class From3D(nn.Module):
    """Runs a 2-D network over every frame of a (N,C,T,h,w) clip and returns (N,C',T,h',w')."""

    def __init__(self, net):
        super().__init__()
        self.model = net

    def forward(self, x):
        N, C, T, h, w = x.shape
        y = self.model(x.permute(0, 2, 1, 3, 4).reshape(N * T, C, h, w))
        return y.view(N, T, *y.shape[-3:]).permute(0, 2, 1, 3, 4)


def make_encoder(args):
    """code/utils/__init__.py:300-351 for the from-scratch model types (pretrained / external checkpoints are out of
    scope: there is no network here and the hot path does not depend on them)."""
    mt = getattr(args, "model_type", "scratch")
    if mt == "scratch":
        net = resnet18()
        net.modify(padding="reflect")
    elif mt == "scratch_zeropad":
        net = resnet18()
    elif mt == "scratch50":
        net = resnet50()
        net.modify(padding="reflect")
    else:
        raise ValueError("model_type %r is not supported by this build (scratch, scratch_zeropad, scratch50)" % mt)
    net.modify(remove_layers=list(getattr(args, "remove_layers", [])))
    return From3D(net)
