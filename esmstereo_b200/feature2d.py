"""2D feature side of ESMStereo (the first "next" row): backbone taps, FPN-style FeatUp,
the image stems and the matching descriptor, with the reference's parameter names
(`models/ESMStereo.py:40-125,528-597`).  Left and right images go through as one batch of 2B.

Two engines, selected per model (`model.feature_engine`):
  * "esm"   -- every conv / transposed conv + BN + activation is one launch of the same fused
               direct-conv kernel the hot path uses (fp32-exact, concat-free); default whenever the
               backbone is the stand-in (`esmstereo_b200.backbone`);
  * "torch" -- plain PyTorch / cuDNN modules; used for a real `timm` backbone's own blocks, whose
               graph this package does not know, and kept for A/B comparisons.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .backbone import FEATURE_CHANS, StandInBackbone, make_backbone
from .layers import _Packed, packed_conv

_TORCH_ACT = {"gelu": F.gelu, "relu": F.relu, "relu6": F.relu6, None: lambda x: x}


def fused(cache: _Packed, conv: nn.Module, bn: Optional[nn.Module], x, act: Optional[str], engine: str, **kw):
    """conv (+BN) (+act) on `x` (tensor or list of channel-concatenated tensors)."""
    if engine in ("esm", "esm_fp32"):  # "esm_fp32": same kernels, tensor-core engines off (see esm_conv_t.engine)
        return ops.conv(x, packed_conv(cache, conv, bn), act, fp32_only=engine == "esm_fp32", **kw)
    if isinstance(x, (list, tuple)):
        x = torch.cat(list(x), 1)
    y = conv(x)
    if bn is not None:
        y = bn(y)
    return _TORCH_ACT[act](y)


class TorchBasicConv(nn.Module):
    """conv(bias=False) -> BN -> GELU with the reference's child names `conv` / `bn` (submodule.py:12-38)."""

    def __init__(self, cin: int, cout: int, deconv: bool = False, **kwargs) -> None:
        super().__init__()
        cls = nn.ConvTranspose2d if deconv else nn.Conv2d
        self.conv = cls(cin, cout, bias=False, **kwargs)
        self.bn = nn.BatchNorm2d(cout)
        self._pc = _Packed()

    def forward(self, x, engine: str = "torch", **kw) -> torch.Tensor:
        return fused(self._pc, self.conv, self.bn, x, "gelu", engine, **kw)


class Conv2x(nn.Module):
    """Reference `Conv2x(deconv=True, concat=True)` (submodule.py:64-103): up-conv, cat skip, 3x3."""

    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        self.conv1 = TorchBasicConv(cin, cout, deconv=True, kernel_size=4, stride=2, padding=1)
        self.conv2 = TorchBasicConv(cout * 2, cout * 2, kernel_size=3, stride=1, padding=1)

    def forward(self, x: torch.Tensor, rem: torch.Tensor, engine: str = "torch") -> torch.Tensor:
        x = self.conv1(x, engine)
        if x.shape != rem.shape:  # never taken for inputs that are multiples of 32 (callers guarantee)
            x = F.interpolate(x.contiguous(), size=(rem.shape[-2], rem.shape[-1]), mode="nearest")
        return self.conv2([x, rem], engine)


class Feature(nn.Module):
    """Backbone taps at 1/2 .. 1/32 (ESMStereo.py:40-77)."""

    def __init__(self, backbone: str) -> None:
        super().__init__()
        self.backbone = backbone
        model = make_backbone(backbone)
        from .timm_compat import TimmCompatBackbone
        self.stand_in = isinstance(model, StandInBackbone)
        self.compat = isinstance(model, TimmCompatBackbone)
        # kernels of libesm_b200 run the stand-in and the timm-compatible definition; a real timm model runs its own forward
        self.esm_capable = self.stand_in or self.compat
        self.chans = FEATURE_CHANS[backbone]
        self.conv_stem, self.bn1, self.act1 = model.conv_stem, model.bn1, nn.ReLU6()
        cuts = [0, 1, 2, 3, 5, 6]
        for i in range(5):
            setattr(self, "block%d" % i, nn.Sequential(*model.blocks[cuts[i]:cuts[i + 1]]))
        self._caches = {}

    def forward(self, x: torch.Tensor, engine: str = "torch") -> List[torch.Tensor]:
        if engine in ("esm", "esm_fp32") and self.stand_in:
            x = fused(self._caches.setdefault("stem", _Packed()), self.conv_stem, self.bn1, x, "relu6", engine)
            outs = []
            for i in range(5):
                for j, stage in enumerate(getattr(self, "block%d" % i)):  # stage = Sequential(conv, bn, ReLU6)
                    x = fused(self._caches.setdefault((i, j), _Packed()), stage[0], stage[1], x, "relu6", engine)
                outs.append(x)
            return outs
        if engine in ("esm", "esm_fp32") and self.compat:
            # `act1(bn1(conv_stem(x)))` (:68): timm's bn1 carries the backbone's activation, act1 = ReLU6 comes on top
            x = fused(self._caches.setdefault("stem", _Packed()), self.conv_stem, self.bn1, x, self.bn1.act_name, engine, act2="relu6")
            outs = []
            for i in range(5):
                for stage in getattr(self, "block%d" % i):
                    for blk in stage:
                        x = blk.forward_esm(x, fp32_only=engine == "esm_fp32")
                outs.append(x)
            return outs
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

    def forward(self, feats: List[torch.Tensor], engine: str = "torch") -> List[torch.Tensor]:
        x2, x4, x8, x16, x32 = feats
        x16 = self.deconv32_16(x32, x16, engine)
        if self.v == 16:
            x16 = self.conv16(x16, engine)
        if self.v in (8, 4):
            x8 = self.deconv16_8(x16, x8, engine)
        if self.v == 8:
            x8 = self.conv8(x8, engine)
        if self.v == 4:
            x4 = self.conv4(self.deconv8_4(x8, x4, engine), engine)
        return [x4, x8, x16, x32]


class ImageStem(nn.Sequential):
    """stem_2/4/8/16 (ESMStereo.py:529-583): BasicConv(s2) + Conv2d + BN + ReLU, children 0..3."""

    def __init__(self, cin: int, cout: int) -> None:
        super().__init__(TorchBasicConv(cin, cout, kernel_size=3, stride=2, padding=1),
                         nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())
        self._pc = _Packed()

    def forward(self, x: torch.Tensor, engine: str = "torch") -> torch.Tensor:
        return fused(self._pc, self[1], self[2], self[0](x, engine), "relu", engine)


def image_stem(cin: int, cout: int) -> ImageStem:
    return ImageStem(cin, cout)


class ConvThenPlain(nn.Sequential):
    """`semantic` (ESMStereo.py:606-618): Sequential(BasicConv, Conv2d(bias=False)), children 0,1."""

    def __init__(self, cin: int, cmid: int, cout: int) -> None:
        super().__init__(TorchBasicConv(cin, cmid, kernel_size=3, stride=1, padding=1), nn.Conv2d(cmid, cout, 3, 1, 1, bias=False))
        self._pc = _Packed()

    def forward(self, x: torch.Tensor, engine: str = "torch") -> torch.Tensor:
        return fused(self._pc, self[1], None, self[0](x, engine), None, engine)
