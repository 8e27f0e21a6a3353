"""Hot-path modules: same parameter tree (names, shapes) as the reference's `models/submodule.py`,
`models/shufflemixer.py`, `models/ESMStereo.py` and `models/ESMStereo_confidence.py`, so reference
checkpoints load by key -- but every `forward` launches the sm_100a kernels of libesm_b200 through
`esmstereo_b200.ops`.  torch.nn layers below are parameter containers only; none of their own
forwards run on this path and there is no PyTorch/CPU fallback.

Inference only: BatchNorm is folded from its running statistics (eval semantics).
"""
from __future__ import annotations

import os
import threading
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops


class _Packed:
    """Lazy, self-invalidating cache of a layer's packed weights, one entry per device (nn.DataParallel replicas share
    the module's __dict__, hence this object: a replica must never be handed weights packed on another GPU).  An entry is
    re-packed when any source tensor is replaced or modified in place (`.cuda()`, `load_state_dict`, optimiser steps).
    Writes through `.data` do not bump a tensor's version: call `invalidate()` (or `esmstereo_b200.layers.invalidate_packed
    (model)`) after editing weights that way."""

    def __init__(self):
        self._entries = {}
        self._lock = threading.Lock()

    def invalidate(self) -> None:
        with self._lock:
            self._entries.clear()

    def get(self, tensors: Sequence[Optional[torch.Tensor]], build):
        live = [t for t in tensors if t is not None]
        dev = live[0].device if live else None
        key = tuple((t.data_ptr(), getattr(t, "_version", 0) if not t.is_inference() else 0) for t in live)
        with self._lock:
            ent = self._entries.get(dev)
            if ent is None or ent[0] != key:
                ent = (key, build())
                self._entries[dev] = ent
            return ent[1]


def invalidate_packed(module: nn.Module) -> int:
    """Drop every packed-weight cache under `module` (after `.data` edits, which the version check cannot see)."""
    n = 0
    for m in module.modules():
        for v in list(vars(m).values()):
            for c in (v.values() if isinstance(v, dict) else v if isinstance(v, (tuple, list)) else (v,)):
                if isinstance(c, _Packed):
                    c.invalidate()
                    n += 1
    return n


def _bn_tuple(bn: Optional[nn.modules.batchnorm._BatchNorm]):
    return None if bn is None else (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


def packed_conv(cache: _Packed, conv: nn.Module, bn: Optional[nn.Module]) -> ops.PackedConv:
    tensors = [conv.weight, conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])
    transposed = isinstance(conv, (nn.ConvTranspose2d, nn.ConvTranspose3d))
    return cache.get(tensors, lambda: ops.pack_conv(conv.weight, conv.stride, conv.padding, transposed, conv.bias,
                                                    _bn_tuple(bn)))


def _all_eq(v, want: int) -> bool:
    return all(a == want for a in v) if isinstance(v, (tuple, list)) else v == want


FUSE_ASSEMBLY = os.environ.get("ESM_FUSE_ASSEMBLY", "0") == "1"  # measured 21 us slower per KITTI pair than the two esm_bilinear_add_f32 launches: opt-in


def _only_full_out_size(fused: dict, x: torch.Tensor) -> bool:
    """True when the only fused argument is an out_size equal to the full 2x output (no crop)."""
    if set(fused) != {"out_size"}:
        return False
    o = fused["out_size"]
    return o is None or tuple(int(v) for v in o) == tuple(2 * int(v) for v in x.shape[2:])


def _inference_only(m: nn.Module) -> None:
    if m.training:
        raise RuntimeError("esmstereo_b200 implements the inference path only (BatchNorm folded from running "
                           "statistics): call model.eval() first")


class BasicConv(nn.Module):
    """Reference `BasicConv` (submodule.py:12-38): conv(bias=False) -> BN -> exact GELU, one kernel.
    Note the reference always instantiates `.bn`, even with bn=False; so do we (state_dict parity)."""

    def __init__(self, in_channels: int, out_channels: int, deconv: bool = False, is_3d: bool = False, bn: bool = True,
                 gelu: bool = True, **kwargs) -> None:
        super().__init__()
        self.gelu, self.use_bn = gelu, bn
        if is_3d:
            cls = nn.ConvTranspose3d if deconv else nn.Conv3d
            self.conv = cls(in_channels, out_channels, bias=False, **kwargs)
            self.bn = nn.BatchNorm3d(out_channels)
        else:
            cls = nn.ConvTranspose2d if deconv else nn.Conv2d
            self.conv = cls(in_channels, out_channels, bias=False, **kwargs)
            self.bn = nn.BatchNorm2d(out_channels)
        self._pc = _Packed()
        self._pc_sub = _Packed()
        self.fp32_only = False  # True: keep this layer off the tensor-core engines (see esm_conv_t.engine)
        # ConvTranspose k4 s2 p1 to ONE channel (conv1_up of `aggregation` and `up_refinement`, ESMStereo.py:150,209):
        # run as its sub-pixel form, a k3 s1 p1 convolution to 2^nd channels (one per output phase) + PixelShuffle.
        # The generic transposed path pads Cout 1 -> 4 and spends 3/4 of its FMAs on zeros (85 us for the cost volume
        # at KITTI shape); the k3 form has no padding waste and is eligible for the tensor-core engine.
        ks = kwargs.get("kernel_size")
        ks = tuple(ks) if isinstance(ks, (tuple, list)) else (ks,) * (3 if is_3d else 2)
        self._subpixel = bool(deconv and out_channels == 1 and all(k == 4 for k in ks) and _all_eq(kwargs.get("stride"), 2)
                              and _all_eq(kwargs.get("padding"), 1))

    def packed(self) -> ops.PackedConv:
        return packed_conv(self._pc, self.conv, self.bn if self.use_bn else None)

    def packed_subpixel(self) -> ops.PackedConv:
        """Phase p = (pd, ph, pw) of the transposed conv reads input offsets p - 1 + t, t in {0, 1}, with kernel taps
        3 - p - 2t: as a k3 p1 conv, tap index p + t of output channel p (channel order = PixelShuffle's)."""
        conv, bn = self.conv, (self.bn if self.use_bn else None)
        tensors = [conv.weight, conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])

        def build():
            w = conv.weight.detach()  # [Cin, 1, 4, 4(, 4)]
            nd = w.dim() - 2
            P = 2 ** nd
            weq = torch.zeros((P, w.shape[0]) + (3,) * nd, device=w.device, dtype=w.dtype)
            for p in range(P):
                ph = [(p >> (nd - 1 - i)) & 1 for i in range(nd)]  # most significant bit = outermost dim
                for t in range(P):
                    tt = [(t >> (nd - 1 - i)) & 1 for i in range(nd)]
                    dst = tuple(ph[i] + tt[i] for i in range(nd))
                    src = tuple(3 - ph[i] - 2 * tt[i] for i in range(nd))
                    weq[(p, slice(None)) + dst] = w[(slice(None), 0) + src]
            rep = lambda t: None if t is None else t.detach().expand(P).contiguous()
            bnt = None if bn is None else (rep(bn.weight), rep(bn.bias), rep(bn.running_mean), rep(bn.running_var), bn.eps)
            return ops.pack_conv(weq, 1, 1, False, rep(conv.bias), bnt)

        return self._pc_sub.get(tensors, build)

    def forward(self, x, **fused) -> torch.Tensor:
        _inference_only(self)
        act = "gelu" if self.gelu else None
        bil = fused.pop("bilinear_prev", None)      # 2D sub-pixel form only: + bilinear x2 of this map, then * final_scale
        final_scale = fused.pop("final_scale", 1.0)  # (the final assembly, ESMStereo.py:307,316, in this layer's epilogue)
        if self._subpixel and isinstance(x, torch.Tensor) and (not fused or set(fused) == {"keep_subpixel"} or _only_full_out_size(fused, x)):
            pc = self.packed_subpixel()
            if x.dim() == 4:
                return ops.conv(x, pc, act, pixel_shuffle=2, fp32_only=self.fp32_only, residual=bil, out_scale=final_scale)
            assert bil is None
            y = ops.conv(x, pc, act, fp32_only=self.fp32_only)  # [B, 8, D, H, W], channel = pd*4 + ph*2 + pw
            if fused.get("keep_subpixel"):
                return y  # the caller reads the phases directly (ops.regression_top2_subpixel)
            return ops.pixel_shuffle3d(y)
        fused.pop("keep_subpixel", None)
        y = ops.conv(x, self.packed(), act, fp32_only=self.fp32_only, **fused)
        return y if bil is None else ops.bilinear_add(bil, y, 2, final_scale)


class ConvBNAct(nn.Sequential):
    """`nn.Sequential(BasicConv, Conv2d(bias=False), BatchNorm2d, act)` -- spx_* (ESMStereo.py:255-258,
    283-285) and conf_spx_4 (ESMStereo_confidence.py:525-528).  Children keep the names 0,1,2,3."""

    def __init__(self, cin: int, cmid: int, cout: int, act: str) -> None:
        super().__init__(BasicConv(cin, cmid, kernel_size=3, stride=1, padding=1),
                         nn.Conv2d(cmid, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout),
                         nn.GELU() if act == "gelu" else nn.ReLU())
        self.act = act
        self._pc = _Packed()

    def forward(self, srcs) -> torch.Tensor:
        _inference_only(self)
        x = self[0](srcs)
        return ops.conv(x, packed_conv(self._pc, self[1], self[2]), self.act)


def bare_conv(cache: _Packed, conv: nn.Module, x, act: Optional[str] = None, **fused) -> torch.Tensor:
    return ops.conv(x, packed_conv(cache, conv, None), act, **fused)


# ---------------------------------------------------------------------------------------------
# 3D hourglass (ESMStereo.py:129-182)
# ---------------------------------------------------------------------------------------------
class aggregation(nn.Module):
    def __init__(self, in_channels: int, add_channel: int) -> None:
        super().__init__()
        c0, c1, c2, c3 = in_channels, in_channels + add_channel, in_channels + add_channel * 2, in_channels + add_channel * 4
        k3 = dict(is_3d=True, kernel_size=3, padding=1)
        self.conv1 = nn.Sequential(BasicConv(c0, c1, stride=2, **k3), BasicConv(c1, c1, stride=1, **k3))
        self.conv2 = nn.Sequential(BasicConv(c1, c2, stride=2, **k3), BasicConv(c2, c2, stride=1, **k3))
        self.conv3 = nn.Sequential(BasicConv(c2, c3, stride=2, **k3), BasicConv(c3, c3, stride=1, **k3))
        up = dict(deconv=True, is_3d=True, kernel_size=(4, 4, 4), padding=(1, 1, 1), stride=(2, 2, 2))
        self.conv3_up = BasicConv(c3, c2, **up)
        self.conv2_up = BasicConv(c2, c1, **up)
        self.conv1_up = BasicConv(c1, 1, bn=False, gelu=False, **up)
        self.agg_0 = nn.Sequential(BasicConv(2 * c2, c2, is_3d=True, kernel_size=1, padding=0, stride=1),
                                   BasicConv(c2, c2, stride=1, **k3))
        self.agg_1 = nn.Sequential(BasicConv(2 * c1, c1, is_3d=True, kernel_size=1, padding=0, stride=1),
                                   BasicConv(c1, c1, stride=1, **k3))

    def forward(self, x: torch.Tensor, keep_subpixel: bool = False) -> torch.Tensor:
        """keep_subpixel: return `conv1_up`'s output as its 8 sub-pixel phases [B,8,D/2,H/2,W/2] instead of the shuffled
        cost volume [B,1,D,H,W] (for ops.regression_top2_subpixel); only honoured when the layer runs in that form."""
        conv1 = self.conv1[1](self.conv1[0](x))
        conv2 = self.conv2[1](self.conv2[0](conv1))
        conv3 = self.conv3[1](self.conv3[0](conv2))
        # crop-to-skip (:172,:177) = only the kept voxels are computed; cat = two-source K loop
        up3 = self.conv3_up(conv3, out_size=conv2.shape[2:])
        conv2 = self.agg_0[1](self.agg_0[0]([up3, conv2]))
        up2 = self.conv2_up(conv2, out_size=conv1.shape[2:])
        conv1 = self.agg_1[1](self.agg_1[0]([up2, conv1]))
        if keep_subpixel and self.conv1_up._subpixel:
            return self.conv1_up(conv1, keep_subpixel=True)
        return self.conv1_up(conv1)


# ---------------------------------------------------------------------------------------------
# ShuffleMixer blocks (shufflemixer.py:23-132)
# ---------------------------------------------------------------------------------------------
class SplitPointMlp(nn.Module):
    def __init__(self, dim: int, mlp_ratio: int = 2) -> None:
        super().__init__()
        hidden = int(dim // 2 * mlp_ratio)
        self.fc = nn.Sequential(nn.Conv2d(dim // 2, hidden, 1, 1, 0), nn.SiLU(inplace=True), nn.Conv2d(hidden, dim // 2, 1, 1, 0))


class BiasFree_LayerNorm(nn.Module):
    def __init__(self, normalized_shape: int) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))


class LayerNorm(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.body = BiasFree_LayerNorm(dim)


class SMLayer(nn.Module):
    """LN -> split-MLP -> shuffle -> +x ; depthwise kxk ; LN -> split-MLP -> shuffle -> +x, as one kernel (k = 7) or two."""

    def __init__(self, dim: int, kernel_size: int, mlp_ratio: int = 2) -> None:
        super().__init__()
        self.norm1, self.norm2 = LayerNorm(dim), LayerNorm(dim)
        self.spatial = nn.Conv2d(dim, dim, kernel_size, 1, kernel_size // 2, groups=dim)
        self.mlp1, self.mlp2 = SplitPointMlp(dim, mlp_ratio), SplitPointMlp(dim, mlp_ratio)
        self._c1, self._c2 = _Packed(), _Packed()
        # One launch per SMLayer (esm_sm_layer_f32: pointwise half on tile + halo, depthwise with one thread per
        # (channel, 4-pixel strip), second pointwise half; bit-identical to the two half kernels).  ESM_SMLAYER=0 runs the
        # two half kernels (esm_sm_pointwise_f32 + esm_sm_spatial_f32), which also cover k != 7.
        self.fused = os.environ.get("ESM_SMLAYER", "1") == "1"

    def _mlp(self, cache: _Packed, norm: LayerNorm, mlp: SplitPointMlp) -> ops.MixerMlp:
        ts = [norm.body.weight, mlp.fc[0].weight, mlp.fc[0].bias, mlp.fc[2].weight, mlp.fc[2].bias]
        return cache.get(ts, lambda: ops.MixerMlp(*ts))

    def forward(self, x: torch.Tensor, extra_residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        _inference_only(self)
        m1, m2 = self._mlp(self._c1, self.norm1, self.mlp1), self._mlp(self._c2, self.norm2, self.mlp2)
        if self.spatial.kernel_size[0] == 7 and self.fused:
            return ops.sm_layer(x, m1, self.spatial.weight.detach(), self.spatial.bias.detach(), m2, extra_residual)
        x = ops.sm_pointwise(x, m1)
        return ops.sm_spatial(x, self.spatial.weight.detach(), self.spatial.bias.detach(), m2, extra_residual)


class FMBlock(nn.Module):
    def __init__(self, dim: int, kernel_size: int, mlp_ratio: int = 2) -> None:
        super().__init__()
        self.net = nn.Sequential(SMLayer(dim, kernel_size, mlp_ratio), SMLayer(dim, kernel_size, mlp_ratio))
        self.conv = nn.Sequential(nn.Conv2d(dim, dim + 16, 3, 1, 1), nn.SiLU(inplace=True), nn.Conv2d(dim + 16, dim, 1, 1, 0))
        self._c0, self._c2 = _Packed(), _Packed()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.net[1](self.net[0](x), extra_residual=x)  # net(x) + x, :130
        y = bare_conv(self._c0, self.conv[0], x, "silu")
        return bare_conv(self._c2, self.conv[2], y, None, residual=x)  # conv(x) + x, :131


# ---------------------------------------------------------------------------------------------
# context-guided disparity upsampling (ESMStereo.py:185-509)
# ---------------------------------------------------------------------------------------------
class up_refinement(nn.Module):
    def __init__(self, C: int, cf1: int, cf2: int) -> None:
        super().__init__()
        k3 = dict(is_3d=False, kernel_size=3, padding=1)
        self.conv1 = nn.Sequential(BasicConv(1, C, stride=2, **k3), BasicConv(C, C, stride=1, **k3))
        self.conv2 = nn.Sequential(BasicConv(C, C, stride=2, **k3), BasicConv(C, C, stride=1, **k3))
        self.conv3 = nn.Sequential(BasicConv(C, C, stride=2, **k3), BasicConv(C, C, stride=1, **k3))
        up = dict(deconv=True, is_3d=False, kernel_size=4, padding=1, stride=2)
        self.conv3_up = BasicConv(C, C, **up)
        self.conv2_up = BasicConv(C, C, **up)
        self.conv1_up = BasicConv(C, 1, bn=False, gelu=False, **up)
        self.agg_0 = nn.Sequential(BasicConv(2 * C + cf1, C, kernel_size=1, padding=0, stride=1), BasicConv(C, C, stride=1, **k3))
        self.agg_1 = nn.Sequential(BasicConv(2 * C + cf2, C, kernel_size=1, padding=0, stride=1), BasicConv(C, C, stride=1, **k3))

    def forward(self, disp, left_f1x, left_f2x, **last_fused) -> torch.Tensor:
        conv1 = self.conv1[1](self.conv1[0](disp))
        conv2 = self.conv2[1](self.conv2[0](conv1))
        conv3 = self.conv3[1](self.conv3[0](conv2))
        up3 = self.conv3_up(conv3, out_size=conv2.shape[2:])  # cropped to conv2, :230
        conv2 = self.agg_0[1](self.agg_0[0]([up3, conv2, left_f1x]))
        up2 = self.conv2_up(conv2)  # NOT cropped, :234
        conv1 = self.agg_1[1](self.agg_1[0]([up2, conv1, left_f2x]))
        return self.conv1_up(conv1, **last_fused)


class UpShuffle(nn.Sequential):
    """`Sequential(Conv2d(n, n*r*r, 1), PixelShuffle(r), SiLU)` as one kernel (ESMStereo.py:265-268)."""

    def __init__(self, n_feats: int, r: int) -> None:
        super().__init__(nn.Conv2d(n_feats, n_feats * r * r, 1, 1, 0), nn.PixelShuffle(r), nn.SiLU(inplace=True))
        self.r = r
        self._pc = _Packed()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return bare_conv(self._pc, self[0], x, "silu", pixel_shuffle=self.r)


def _disp_mlp(C: int) -> nn.Sequential:
    """dm2x/dm4x/dm8x/cm: k5 p1, k3, k3, k1 p1 (ESMStereo.py:250-253)."""
    return nn.Sequential(BasicConv(1, C, is_3d=False, kernel_size=5, padding=1, stride=1),
                         BasicConv(C, C, is_3d=False, kernel_size=3, padding=1, stride=1),
                         BasicConv(C, C, is_3d=False, kernel_size=3, padding=1, stride=1),
                         BasicConv(C, C, is_3d=False, kernel_size=1, padding=1, stride=1))


def _run_seq(seq: nn.Sequential, x):
    for m in seq:
        x = m(x)
    return x


class _Upsampler(nn.Module):
    """Shared body of upsample4 / upsample8 / upsample16.  `stages` = [(tag, C, cf1, cf2, Cspx_out)]."""

    def _build(self, stages, n_feats: int, r: int) -> None:
        self.r = r
        self._tags = [s[0] for s in stages]
        self._ct = {}
        for i, (tag, Cc, cf1, cf2, cout) in enumerate(stages):
            setattr(self, "dm" + tag, _disp_mlp(Cc))
            setattr(self, "spx_" + tag, ConvBNAct(Cc + cf2, Cc, cout, "gelu"))
            if i == 0:
                self.to_feat = nn.Conv2d(Cc, n_feats, 3, 1, 1, bias=False)
                self.blocks = nn.Sequential(*[FMBlock(n_feats, 7, 2) for _ in range(2)])
                self._tf = _Packed()
            setattr(self, "upsampling" + tag[0], UpShuffle(n_feats, r))
            setattr(self, "tail" + tag, nn.Conv2d(n_feats, 1, 3, 1, 1))
            setattr(self, "ref" + tag, up_refinement(Cc, cf1, cf2))
            self._ct[tag] = _Packed()

    def _stage(self, i: int, prev, spx_feat, ref_f1, ref_f2, final_scale: float = 1.0):
        tag = self._tags[i]
        x = _run_seq(getattr(self, "dm" + tag), prev)
        x = getattr(self, "spx_" + tag)([x, spx_feat])
        if i == 0:
            x = bare_conv(self._tf, self.to_feat, x)
            x = _run_seq(self.blocks, x)
        x = getattr(self, "upsampling" + tag[0])(x)
        x = bare_conv(self._ct[tag], getattr(self, "tail" + tag), x)
        if self.r == 2 and FUSE_ASSEMBLY:  # upsampled + refined, * final_scale, in conv1_up's epilogue
            return getattr(self, "ref" + tag)(x, ref_f1, ref_f2, bilinear_prev=prev, final_scale=final_scale)
        x = getattr(self, "ref" + tag)(x, ref_f1, ref_f2)
        return ops.bilinear_add(prev, x, self.r, final_scale)


class upsample4(_Upsampler):
    def __init__(self) -> None:
        super().__init__()
        #            tag   C  cf1 cf2 spx_out
        self._build([("2x", 32, 96, 48, 32), ("4x", 32, 48, 32, 16)], n_feats=16, r=2)

    def forward(self, left_f1x, left_f2x, left_f4x, init_disp, out_scale: float = 1.0):
        d2 = self._stage(0, init_disp, left_f2x, left_f1x, left_f2x)
        d4 = self._stage(1, d2, left_f4x, left_f2x, left_f4x, out_scale)
        return d4, d2


class upsample8(_Upsampler):
    def __init__(self) -> None:
        super().__init__()
        self._build([("2x", 16, 240, 96, 16), ("4x", 16, 96, 24, 8), ("8x", 16, 24, 32, 8)], n_feats=8, r=2)

    def forward(self, left_f2x, left_f4x, left_f8x, stem_f2, init_disp, out_scale: float = 1.0):
        d2 = self._stage(0, init_disp, left_f4x, left_f2x, left_f4x)
        d4 = self._stage(1, d2, left_f8x, left_f4x, left_f8x)
        d8 = self._stage(2, d4, stem_f2, left_f8x, stem_f2, out_scale)
        return d8, d4, d2


class upsample16(_Upsampler):
    def __init__(self) -> None:
        super().__init__()
        self._build([("2x", 16, 32, 32, 16), ("4x", 16, 24, 24, 8)], n_feats=8, r=4)

    def forward(self, left_f1x, left_f2x, left_f4x, left_f8x, init_disp, out_scale: float = 1.0):
        d2 = self._stage(0, init_disp, left_f2x, left_f2x, left_f1x)
        d4 = self._stage(1, d2, left_f4x, left_f4x, left_f8x, out_scale)
        return d4, d2


# ---------------------------------------------------------------------------------------------
# confidence head (ESMStereo_confidence.py:511-744)
# ---------------------------------------------------------------------------------------------
class conf_upsample(nn.Module):
    def __init__(self, C: int, fc: int) -> None:
        super().__init__()
        self.conv1 = BasicConv(1, C, is_3d=False, kernel_size=3, padding=1, stride=1, dilation=1)
        self.conv2 = BasicConv(C, C, is_3d=False, kernel_size=3, padding=1, stride=2, dilation=1)
        self.conv1_up = BasicConv(C, 1, deconv=True, is_3d=False, kernel_size=4, padding=1, stride=2)
        self.cm = _disp_mlp(C)
        self.conf_spx_4 = ConvBNAct(C + fc, C, C, "relu")
        self.conf_spx = nn.ConvTranspose2d(C, 9, kernel_size=4, stride=4, padding=0)

    def forward(self, left_f1x, init_conf, final_act: Optional[str] = None) -> torch.Tensor:
        x = _run_seq(self.cm, init_conf)
        x = self.conf_spx_4([x, left_f1x])
        conf1 = ops.conf_convex_up4(x, init_conf, self.conf_spx.weight.detach().contiguous(), self.conf_spx.bias.detach())
        y = self.conv2(self.conv1(conf1))
        return self.conv1_up(y, residual=conf1, act2=final_act)  # conf + conf1 (:548) [+ sigmoid :744]


class LAFNet_ESM(nn.Module):
    def __init__(self, C: int) -> None:
        super().__init__()
        self.C = C
        for name, cin in (("cost", 7), ("disp", 1), ("imag", 64)):
            setattr(self, name + "_conv1", nn.Conv2d(cin, C, kernel_size=3, padding=1))
            setattr(self, name + "_bn1", nn.BatchNorm2d(C))
            setattr(self, name + "_conv2", nn.Conv2d(C, C, kernel_size=3, padding=1))
            setattr(self, name + "_bn2", nn.BatchNorm2d(C))
            setattr(self, name + "_conv3", nn.Conv2d(C, C, kernel_size=1, padding=0))
            setattr(self, name + "_bn3", nn.BatchNorm2d(C))
        for name in ("cost", "disp", "imag"):
            setattr(self, name + "_att_conv1", nn.Conv2d(C, C, kernel_size=3, padding=1))
            setattr(self, name + "_att_bn1", nn.BatchNorm2d(C))
            setattr(self, name + "_att_conv2", nn.Conv2d(C, 1, kernel_size=1, padding=0))
            setattr(self, name + "_att_bn2", nn.BatchNorm2d(1))
        self.scale_conv1, self.scale_bn1 = nn.Conv2d(C, C, kernel_size=3, padding=1), nn.BatchNorm2d(C)
        self.scale_conv2, self.scale_bn2 = nn.Conv2d(C, C, kernel_size=3, padding=1), nn.BatchNorm2d(C)
        self.scale_conv3, self.scale_bn3 = nn.Conv2d(C, 1, kernel_size=1, padding=0), nn.BatchNorm2d(1)
        self.embed_conv1, self.embed_bn1 = nn.Conv2d(3 * C, C, kernel_size=3, padding=1), nn.BatchNorm2d(C)
        self.embed_conv2, self.embed_bn2 = nn.Conv2d(C, C, kernel_size=3, padding=0, stride=3), nn.BatchNorm2d(C)
        self.fusion_conv1 = nn.Conv2d(C + 1, C, kernel_size=3, padding=1)
        self.fusion_conv2 = nn.Conv2d(C, C, kernel_size=3, padding=1)
        self.fusion_conv3 = nn.Conv2d(C, 1, kernel_size=1, padding=0)
        for i in (1, 2, 3):
            for j, ch in ((1, C), (2, C), (3, 1)):
                setattr(self, "fusion_bn%d_iter%d" % (j, i), nn.BatchNorm2d(ch))
        self.conf_up4 = conf_upsample(C, 96)
        self.conf_up1 = conf_upsample(C, 24)
        nn.init.constant_(self.scale_bn3.weight, 0)  # ":641-642"
        nn.init.constant_(self.scale_bn3.bias, 0)
        self._caches = {}
        self.capture = None  # set to {} to record stages (tests)

    def _cb(self, conv: str, bn: str, x, act: Optional[str] = "relu", **fused):
        cache = self._caches.setdefault((conv, bn), _Packed())
        pc = packed_conv(cache, getattr(self, conv), getattr(self, bn))
        return ops.conv(x, pc, act, **fused)

    def forward(self, cost, disp, imag, left_f1x, left_f2x, device=None) -> torch.Tensor:
        _inference_only(self)
        tower = {}
        for name, src in (("cost", ops.laf_cost_top7(cost)), ("disp", disp), ("imag", imag)):
            t = self._cb(name + "_conv1", name + "_bn1", src)
            t = self._cb(name + "_conv2", name + "_bn2", t)
            tower[name] = self._cb(name + "_conv3", name + "_bn3", t)
        att = []
        for name in ("cost", "disp", "imag"):
            t = self._cb(name + "_att_conv1", name + "_att_bn1", tower[name])
            att.append(self._cb(name + "_att_conv2", name + "_att_bn2", t, None))
        x = ops.laf_attention(tower["cost"], tower["disp"], tower["imag"], *att)
        feat = self._cb("embed_conv1", "embed_bn1", x)
        t = self._cb("scale_conv1", "scale_bn1", feat)
        t = self._cb("scale_conv2", "scale_bn2", t)
        scale = self._cb("scale_conv3", "scale_bn3", t, "2sigmoid")
        # scale-adaptive sampling + embed_conv2 (k3 s3) + BN + ReLU, fused (":693-719")
        cache = self._caches.setdefault(("embed2",), _Packed())
        bn = self.embed_bn2
        ts = [self.embed_conv2.weight, self.embed_conv2.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]

        def build():
            w = self.embed_conv2.weight.detach().contiguous()
            sc = torch.empty(self.C, device=w.device)
            sh = torch.empty(self.C, device=w.device)
            ops.check(ops.lib().esm_fold_bn_f32(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                                bn.running_var.data_ptr(), self.embed_conv2.bias.data_ptr(), float(bn.eps),
                                                self.C, sc.data_ptr(), sh.data_ptr(), ops._stream()), "fold_bn")
            return w, sc, sh

        w, sc, sh = cache.get(ts, build)
        if self.capture is not None:
            self.capture.update(conf_top7=ops.laf_cost_top7(cost), conf_feat=feat, conf_scale=scale)
        feat = ops.laf_sample_embed(feat, scale, w, sc, sh)
        out = ops.fill_(torch.empty_like(disp), 0.5)
        for it in (1, 2, 3):  # shared convs, per-iteration BN (":725-739")
            t = self._cb("fusion_conv1", "fusion_bn1_iter%d" % it, [feat, out])
            t = self._cb("fusion_conv2", "fusion_bn2_iter%d" % it, t)
            out = self._cb("fusion_conv3", "fusion_bn3_iter%d" % it, t)
        out4 = self.conf_up4(left_f1x, out)
        if self.capture is not None:
            self.capture.update(conf_embed=feat, conf_init=out, conf_4=out4)
        return self.conf_up1(left_f2x, out4, final_act="sigmoid")
