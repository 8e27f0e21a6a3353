"""timm-key-compatible backbones for `Feature` (`/root/reference/models/ESMStereo.py:40-77`).

The reference builds its backbone with `timm.create_model('efficientnet_b2' | 'mobilenetv2_100', pretrained=True,
features_only=True)` and keeps `conv_stem`, `bn1` and `blocks[0:6]`, so its checkpoints (`esmstereo_{L,S}_gwc.ckpt`)
carry the backbone under timm's parameter names:

    feature.conv_stem.weight, feature.bn1.*,
    feature.block{0..4}.{stage}.{block}.conv_dw / bn1 / se.conv_reduce / se.conv_expand / conv_pw / bn2        (DepthwiseSeparableConv)
    feature.block{0..4}.{stage}.{block}.conv_pw / bn1 / conv_dw / bn2 / se.* / conv_pwl / bn3                  (InvertedResidual)

timm is not installed in this image, so this module restates the two architectures -- EfficientNet-B2 (B0's block
table with width x1.1 / depth x1.2, SiLU, squeeze-and-excitation at 0.25 of the block input) and MobileNetV2-1.0
(ReLU6, no SE) -- with exactly those module and parameter names and shapes, so that the reference's key-filtered
`load_state_dict` (test_kitti.py:57-61) fills every backbone tensor.  Two execution paths per block: the modules' own
PyTorch forward ("torch": cuDNN) and `forward_esm` (libesm_b200: 1x1 convolutions on the conv engines, depthwise
conv + BN + activation, global pooling and the SE gate as CUDA kernels of csrc/backbone.cu).

Parity of these definitions against a real timm cannot be checked here (no timm, no network, no checkpoints): the key
names and shapes follow timm's `efficientnet_b2` / `mobilenetv2_100` as of timm 0.9/1.0 (BatchNormAct2d: the BN
modules carry their activation and have plain BatchNorm2d state), and tests pin them against the expected key list.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layers import _Packed

_ACT = {"silu": F.silu, "relu6": F.relu6, None: lambda x: x}


class BatchNormAct2d(nn.BatchNorm2d):
    """BatchNorm2d that applies its activation, like timm's (state_dict identical to nn.BatchNorm2d)."""

    def __init__(self, c: int, act: Optional[str]) -> None:
        super().__init__(c)
        self.act_name = act

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _ACT[self.act_name](super().forward(x))


def _fold(bn: nn.BatchNorm2d):
    sc = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    return sc.contiguous(), (bn.bias.detach() - bn.running_mean.detach() * sc).contiguous()


class SqueezeExcite(nn.Module):
    def __init__(self, chs: int, rd: int, act: str) -> None:
        super().__init__()
        self.conv_reduce = nn.Conv2d(chs, rd, 1, bias=True)
        self.conv_expand = nn.Conv2d(rd, chs, 1, bias=True)
        self.act_name = act
        self._c = (_Packed(), _Packed())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(_ACT[self.act_name](self.conv_reduce(s)))
        return x * torch.sigmoid(s)

    def forward_esm(self, x: torch.Tensor) -> torch.Tensor:
        from .layers import bare_conv
        s = ops.global_avgpool(x)
        s = bare_conv(self._c[0], self.conv_reduce, s, self.act_name, fp32_only=True)
        g = bare_conv(self._c[1], self.conv_expand, s, "sigmoid", fp32_only=True)
        return ops.scale_channels_(x, g)


class _Block(nn.Module):
    def _dw(self, x, conv, bn, act):
        sc, sh = self._fold_cache(bn)
        return ops.dwconv2d(x, conv.weight.detach(), sc, sh, act, conv.stride[0])

    def _fold_cache(self, bn):
        c = self.__dict__.setdefault("_folds", {})
        key = id(bn)
        ent = c.get(key)
        ver = tuple((t.data_ptr(), t._version) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))
        if ent is None or ent[0] != ver:
            ent = (ver, _fold(bn))
            c[key] = ent
        return ent[1]


class DepthwiseSeparableConv(_Block):
    """timm `DepthwiseSeparableConv`: dw k x k -> BN+act -> [SE] -> pw 1x1 -> BN (no act), + skip when shapes allow."""

    def __init__(self, cin: int, cout: int, k: int, stride: int, act: str, se_rd: int = 0) -> None:
        super().__init__()
        self.has_skip = stride == 1 and cin == cout
        self.conv_dw = nn.Conv2d(cin, cin, k, stride, k // 2, groups=cin, bias=False)
        self.bn1 = BatchNormAct2d(cin, act)
        self.se = SqueezeExcite(cin, se_rd, act) if se_rd else nn.Identity()
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = BatchNormAct2d(cout, None)
        self._pc = _Packed()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.bn2(self.conv_pw(self.se(self.bn1(self.conv_dw(x)))))
        return y + x if self.has_skip else y

    def forward_esm(self, x: torch.Tensor, fp32_only: bool = False) -> torch.Tensor:
        from .layers import packed_conv
        y = self._dw(x, self.conv_dw, self.bn1, self.bn1.act_name)
        if isinstance(self.se, SqueezeExcite):
            y = self.se.forward_esm(y)
        return ops.conv(y, packed_conv(self._pc, self.conv_pw, self.bn2), None, residual=x if self.has_skip else None, fp32_only=fp32_only)


class InvertedResidual(_Block):
    """timm `InvertedResidual` (MBConv): pw 1x1 -> BN+act -> dw k x k -> BN+act -> [SE] -> pwl 1x1 -> BN, + skip."""

    def __init__(self, cin: int, cout: int, k: int, stride: int, exp: int, act: str, se_rd: int = 0) -> None:
        super().__init__()
        mid = cin * exp
        self.has_skip = stride == 1 and cin == cout
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = BatchNormAct2d(mid, act)
        self.conv_dw = nn.Conv2d(mid, mid, k, stride, k // 2, groups=mid, bias=False)
        self.bn2 = BatchNormAct2d(mid, act)
        self.se = SqueezeExcite(mid, se_rd, act) if se_rd else nn.Identity()
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = BatchNormAct2d(cout, None)
        self._pc = (_Packed(), _Packed())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.bn1(self.conv_pw(x))
        y = self.se(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(y))
        return y + x if self.has_skip else y

    def forward_esm(self, x: torch.Tensor, fp32_only: bool = False) -> torch.Tensor:
        from .layers import packed_conv
        y = ops.conv(x, packed_conv(self._pc[0], self.conv_pw, self.bn1), self.bn1.act_name, fp32_only=fp32_only)
        y = self._dw(y, self.conv_dw, self.bn2, self.bn2.act_name)
        if isinstance(self.se, SqueezeExcite):
            y = self.se.forward_esm(y)
        return ops.conv(y, packed_conv(self._pc[1], self.conv_pwl, self.bn3), None, residual=x if self.has_skip else None, fp32_only=fp32_only)


def _round_channels(c: float, divisor: int = 8) -> int:
    new = max(divisor, int(c + divisor / 2) // divisor * divisor)
    if new < 0.9 * c:  # timm / EfficientNet: never round down by more than 10 %
        new += divisor
    return new


# (block type, repeats, kernel, stride, expansion, channels) of the 7 stages -- EfficientNet-B0 / MobileNetV2 tables
_EFFNET_B0 = [("ds", 1, 3, 1, 1, 16), ("ir", 2, 3, 2, 6, 24), ("ir", 2, 5, 2, 6, 40), ("ir", 3, 3, 2, 6, 80), ("ir", 3, 5, 1, 6, 112),
              ("ir", 4, 5, 2, 6, 192), ("ir", 1, 3, 1, 6, 320)]
_MBV2 = [("ds", 1, 3, 1, 1, 16), ("ir", 2, 3, 2, 6, 24), ("ir", 3, 3, 2, 6, 32), ("ir", 4, 3, 2, 6, 64), ("ir", 3, 3, 1, 6, 96),
         ("ir", 3, 3, 2, 6, 160), ("ir", 1, 3, 1, 6, 320)]
ARCH = {
    # name: (table, width multiplier, depth multiplier, activation, SE ratio of the block input)
    "efficientnet_b2": (_EFFNET_B0, 1.1, 1.2, "silu", 0.25),
    "mobilenetv2_100": (_MBV2, 1.0, 1.0, "relu6", 0.0),
}


class TimmCompatBackbone(nn.Module):
    """`conv_stem`, `bn1`, `blocks` (7 stages) with timm's names, shapes and arithmetic."""

    def __init__(self, name: str) -> None:
        super().__init__()
        if name not in ARCH:
            raise ValueError("unknown backbone %r" % (name,))
        table, wm, dm, act, se = ARCH[name]
        stem = _round_channels(32 * wm)
        self.conv_stem = nn.Conv2d(3, stem, 3, 2, 1, bias=False)
        self.bn1 = BatchNormAct2d(stem, act)
        stages, cin = [], stem
        for kind, reps, k, stride, exp, ch in table:
            cout = _round_channels(ch * wm)
            blocks = []
            for i in range(int(math.ceil(reps * dm))):
                s = stride if i == 0 else 1
                if kind == "ds":
                    blocks.append(DepthwiseSeparableConv(cin, cout, k, s, act, se_rd=max(1, round(cin * se)) if se else 0))
                else:
                    blocks.append(InvertedResidual(cin, cout, k, s, exp, act, se_rd=max(1, round(cin * se)) if se else 0))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)


def create_model(name: str, pretrained: bool = False, features_only: bool = True, **_unused) -> TimmCompatBackbone:
    """Signature of `timm.create_model` as the reference calls it (no pretrained weights: there is no network)."""
    return TimmCompatBackbone(name)


def stage_channels(name: str) -> List[int]:
    table, wm = ARCH[name][0], ARCH[name][1]
    return [_round_channels(ch * wm) for _k, _r, _kk, _s, _e, ch in table]
