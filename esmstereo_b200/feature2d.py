"""2D feature side of ESMStereo -- NOT the hot path ("kept in PyTorch, must exist for
drop-in").  Backbone taps, FPN-style FeatUp, the image stems and the matching descriptor run as
ordinary PyTorch/cuDNN modules with the reference's parameter names (`models/ESMStereo.py:40-125,
528-597`); left and right images are pushed through as one batch of 2B.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .backbone import FEATURE_CHANS, make_backbone


class TorchBasicConv(nn.Module):
    """conv(bias=False) -> BN -> GELU with the reference's child names `conv` / `bn` (submodule.py:12-38)."""

    def __init__(self, cin: int, cout: int, deconv: bool = False, **kwargs) -> None:
        super().__init__()
        cls = nn.ConvTranspose2d if deconv else nn.Conv2d
        self.conv = cls(cin, cout, bias=False, **kwargs)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.gelu(self.bn(self.conv(x)))


class Conv2x(nn.Module):
    """Reference `Conv2x(deconv=True, concat=True)` (submodule.py:64-103): up-conv, cat skip, 3x3."""

    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        self.conv1 = TorchBasicConv(cin, cout, deconv=True, kernel_size=4, stride=2, padding=1)
        self.conv2 = TorchBasicConv(cout * 2, cout * 2, kernel_size=3, stride=1, padding=1)

    def forward(self, x: torch.Tensor, rem: torch.Tensor) -> torch.Tensor:
        x = self.conv1(x)
        if x.shape != rem.shape:
            x = F.interpolate(x, size=(rem.shape[-2], rem.shape[-1]), mode="nearest")
        return self.conv2(torch.cat((x, rem), 1))


class Feature(nn.Module):
    """Backbone taps at 1/2 .. 1/32 (ESMStereo.py:40-77)."""

    def __init__(self, backbone: str) -> None:
        super().__init__()
        self.backbone = backbone
        model = make_backbone(backbone)
        self.chans = FEATURE_CHANS[backbone]
        self.conv_stem, self.bn1, self.act1 = model.conv_stem, model.bn1, nn.ReLU6()
        cuts = [0, 1, 2, 3, 5, 6]
        for i in range(5):
            setattr(self, "block%d" % i, nn.Sequential(*model.blocks[cuts[i]:cuts[i + 1]]))

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        x = self.act1(self.bn1(self.conv_stem(x)))
        outs = []
        for i in range(5):
            x = getattr(self, "block%d" % i)(x)
            outs.append(x)
        return outs  # x2, x4, x8, x16, x32


class FeatUp(nn.Module):
    """Top-down feature fusion (ESMStereo.py:79-125); one image batch at a time."""

    def __init__(self, chans: List[int], vol_size: int) -> None:
        super().__init__()
        self.v = vol_size
        self.deconv32_16 = Conv2x(chans[4], chans[3])
        if vol_size == 16:
            self.conv16 = TorchBasicConv(chans[3] * 2, chans[2] * 2, kernel_size=3, stride=1, padding=1)
        if vol_size in (8, 4):
            self.deconv16_8 = Conv2x(chans[3] * 2, chans[2])
        if vol_size == 8:
            self.conv8 = TorchBasicConv(chans[2] * 2, chans[2] * 2, kernel_size=3, stride=1, padding=1)
        if vol_size == 4:
            self.deconv8_4 = Conv2x(chans[2] * 2, chans[1])
            self.conv4 = TorchBasicConv(chans[1] * 2, chans[1] * 2, kernel_size=3, stride=1, padding=1)

    def forward(self, feats: List[torch.Tensor]) -> List[torch.Tensor]:
        x2, x4, x8, x16, x32 = feats
        x16 = self.deconv32_16(x32, x16)
        if self.v == 16:
            x16 = self.conv16(x16)
        if self.v in (8, 4):
            x8 = self.deconv16_8(x16, x8)
        if self.v == 8:
            x8 = self.conv8(x8)
        if self.v == 4:
            x4 = self.conv4(self.deconv8_4(x8, x4))
        return [x4, x8, x16, x32]


def image_stem(cin: int, cout: int) -> nn.Sequential:
    """stem_2/4/8/16 (ESMStereo.py:529-583): BasicConv(s2) + Conv2d + BN + ReLU, children 0..3."""
    return nn.Sequential(TorchBasicConv(cin, cout, kernel_size=3, stride=2, padding=1),
                         nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())
