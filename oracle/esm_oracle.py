"""CPU oracle for the ESMStereo feature-to-disparity path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (state_dict-driven) restatement, on torch-CPU tensors, of the algorithm in
`/root/reference/models/{submodule,shufflemixer,ESMStereo,ESMStereo_confidence}.py`.  Only `tests/`,
`__graft_entry__.smoke()` and the CPU-baseline / `--impl reference` legs of `bench.py` may import it;
the product package `esmstereo_b200` never does (it fails loudly without its CUDA library).

Pinning: the reference has no tests, golden vectors or fixtures for this path, so
parity is pinned the only way available -- `tests/golden/make_golden.py` imports the real reference
in the build container (with a `timm` shim), runs it on seeded inputs with name-keyed deterministic
weights, and commits its stage outputs under `tests/golden/`; `tests/test_oracle_golden.py` holds this
oracle to those vectors.  Backbone: only the stand-in backbone of `esmstereo_b200.backbone` is restated
(real timm weights/graphs are absent from the image).

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm{2,3}d default, used everywhere in the reference


class EsmOracle:
    """`EsmOracle(sd, ...)(left, right)` -> dict of stage tensors (see `forward`)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], maxdisp: int = 192, gwc: bool = False,
                 norm_correlation: bool = True, backbone: str = "efficientnet_b2", cv_scale: int = 4,
                 confidence: bool = False, dtype: torch.dtype = torch.float32, device="cpu") -> None:
        # device: "cpu" for the oracle proper; bench.py's `gpu_eager_baseline` runs the same torch calls on "cuda" to time
        # the reference's eager ATen / cuDNN path on the box's GPU (a baseline, never the product)
        self.dtype, self.device = dtype, torch.device(device)
        self.sd = {k: (v.detach().to(self.device).to(dtype) if v.is_floating_point() else v.detach().to(self.device))
                   for k, v in state_dict.items()}
        self.maxdisp, self.gwc, self.ncorr = maxdisp, gwc, norm_correlation
        self.backbone, self.s, self.confidence = backbone, cv_scale, confidence
        self.calibrating = False

    # ------------------------------------------------------------------ primitives
    def w(self, key: str) -> torch.Tensor:
        return self.sd[key]

    def has(self, key: str) -> bool:
        return key in self.sd

    def bn(self, x: torch.Tensor, prefix: str) -> torch.Tensor:
        """Eval-mode BatchNorm (`submodule.py:24,30,35`).  In calibration mode it behaves like a
        train-mode BN with momentum 1.0: normalise with batch statistics and overwrite the running
        statistics (unbiased variance, as torch does) -- SURVEY.md section 7 hard part 3."""
        g, b = self.w(prefix + ".weight"), self.w(prefix + ".bias")
        if self.calibrating:
            dims = [0] + list(range(2, x.dim()))
            mean = x.mean(dims)
            var_b = x.var(dims, unbiased=False)
            n = x.numel() // x.shape[1]
            self.sd[prefix + ".running_mean"] = mean.clone()
            self.sd[prefix + ".running_var"] = (var_b * (n / max(n - 1, 1))).clone()
            shape = [1, -1] + [1] * (x.dim() - 2)
            return (x - mean.view(shape)) / torch.sqrt(var_b.view(shape) + BN_EPS) * g.view(shape) + b.view(shape)
        return F.batch_norm(x, self.w(prefix + ".running_mean"), self.w(prefix + ".running_var"),
                            g, b, False, 0.0, BN_EPS)

    @staticmethod
    def act(x: torch.Tensor, kind: Optional[str]) -> torch.Tensor:
        if kind is None:
            return x
        if kind == "gelu":  # exact erf GELU, `submodule.py:37`
            return F.gelu(x)
        if kind == "relu":
            return F.relu(x)
        if kind == "relu6":
            return F.relu6(x)
        if kind == "silu":
            return F.silu(x)
        raise ValueError(kind)

    def conv(self, x: torch.Tensor, wkey: str, stride=1, pad=1, deconv: bool = False,
             bias: Optional[str] = None, groups: int = 1) -> torch.Tensor:
        wt = self.w(wkey)
        bt = self.w(bias) if bias is not None else None
        if wt.dim() == 5:
            fn = F.conv_transpose3d if deconv else F.conv3d
        else:
            fn = F.conv_transpose2d if deconv else F.conv2d
        return fn(x, wt, bt, stride=stride, padding=pad, groups=groups)

    def basic(self, x: torch.Tensor, prefix: str, stride=1, pad=1, deconv: bool = False,
              bn: bool = True, act: Optional[str] = "gelu") -> torch.Tensor:
        """`BasicConv.forward` (`submodule.py:32-38`): conv(bias=False) -> BN -> GELU."""
        x = self.conv(x, prefix + ".conv.weight", stride, pad, deconv)
        if bn:
            x = self.bn(x, prefix + ".bn")
        return self.act(x, act)

    def conv_bn_act_seq(self, x: torch.Tensor, prefix: str, act: str) -> torch.Tensor:
        """`nn.Sequential(BasicConv(k3), Conv2d(k3,bias=False), BatchNorm2d, act)` used by the stems
        (`ESMStereo.py:529-583`), `spx_*` (`:255-258,283-285`) and `conf_spx_4` (`_confidence.py:525-528`)."""
        x = self.basic(x, prefix + ".0", 1, 1)
        x = self.conv(x, prefix + ".1.weight", 1, 1)
        return self.act(self.bn(x, prefix + ".2"), act)

    # ------------------------------------------------------------------ 2D feature side (not hot path)
    def feature(self, img: torch.Tensor) -> List[torch.Tensor]:
        """`Feature.forward` (`ESMStereo.py:68-77`) over the stand-in backbone."""
        x = F.relu6(self.bn(self.conv(img, "feature.conv_stem.weight", 2, 1), "feature.bn1"))
        stage_strides = {"efficientnet_b2": [1, 2, 2, 2, 1, 2, 1], "mobilenetv2_100": [1, 2, 2, 2, 1, 2, 1]}[self.backbone]
        outs, si = [], 0
        for blk, nstage in enumerate([1, 1, 1, 2, 1]):  # blocks[0:1],[1:2],[2:3],[3:5],[5:6]
            for j in range(nstage):
                p = "feature.block%d.%d" % (blk, j)
                x = F.relu6(self.bn(self.conv(x, p + ".0.weight", stage_strides[si], 1), p + ".1"))
                si += 1
            outs.append(x)
        return outs  # x2, x4, x8, x16, x32

    def conv2x(self, x: torch.Tensor, rem: torch.Tensor, prefix: str) -> torch.Tensor:
        """`Conv2x.forward` with deconv=True, concat=True (`submodule.py:91-103`)."""
        x = self.basic(x, prefix + ".conv1", 2, 1, deconv=True)
        if x.shape != rem.shape:
            x = F.interpolate(x, size=(rem.shape[-2], rem.shape[-1]), mode="nearest")
        return self.basic(torch.cat((x, rem), 1), prefix + ".conv2", 1, 1)

    def feat_up(self, f: List[torch.Tensor]) -> List[torch.Tensor]:
        """`FeatUp.forward` (`ESMStereo.py:100-125`), one image at a time."""
        x2, x4, x8, x16, x32 = f
        x16 = self.conv2x(x32, x16, "feature_up.deconv32_16")
        if self.s in (8, 4):
            x8 = self.conv2x(x16, x8, "feature_up.deconv16_8")
        if self.s == 8:
            x8 = self.basic(x8, "feature_up.conv8")
        if self.s == 4:
            x4 = self.conv2x(x8, x4, "feature_up.deconv8_4")
            x4 = self.basic(x4, "feature_up.conv4")
        return [x4, x8, x16, x32]

    def stem(self, x: torch.Tensor, prefix: str) -> torch.Tensor:
        x = self.basic(x, prefix + ".0", 2, 1)
        x = self.conv(x, prefix + ".1.weight", 1, 1)
        return F.relu(self.bn(x, prefix + ".2"))

    def descriptors(self, img: torch.Tensor):
        """Left or right branch of `ESMStereo.forward` up to `match_*` (`ESMStereo.py:640-695`)."""
        raw = self.feature(img)
        feats = self.feat_up(raw) if self.s in (4, 8) else raw
        stems = [self.stem(img, "stem_2")]
        for name in {4: ["stem_4"], 8: ["stem_4", "stem_8"], 16: ["stem_4", "stem_8", "stem_16"]}[self.s]:
            stems.append(self.stem(stems[-1], name))
        coarse = {4: feats[0], 8: feats[1], 16: feats[3]}[self.s]
        m = torch.cat((coarse, stems[-1]), 1)
        m = self.basic(m, "conv", 1, 1)
        m = self.conv(m, "desc.weight", 1, 0, bias="desc.bias")
        return feats, stems, m

    # ------------------------------------------------------------------ hot path: volumes
    def gwc_volume(self, L: torch.Tensor, R: torch.Tensor, D: int, G: int) -> torch.Tensor:
        """`build_gwc_volume` + `groupwise_correlation` (`submodule.py:143-161`)."""
        B, C, H, W = L.shape
        V = L.new_zeros(B, G, D, H, W)
        for d in range(min(D, W)):
            prod = L[:, :, :, d:] * R[:, :, :, : W - d]
            V[:, :, d, :, d:] = prod.view(B, G, C // G, H, W - d).mean(2)
        return V

    def norm_corr_volume(self, L: torch.Tensor, R: torch.Tensor, D: int) -> torch.Tensor:
        """`build_norm_correlation_volume` + `norm_correlation` (`submodule.py:187-200`)."""
        B, C, H, W = L.shape
        Ln = L / (torch.norm(L, 2, 1, True) + 1e-5)
        Rn = R / (torch.norm(R, 2, 1, True) + 1e-5)
        V = L.new_zeros(B, 1, D, H, W)
        for d in range(min(D, W)):
            V[:, :, d, :, d:] = (Ln[:, :, :, d:] * Rn[:, :, :, : W - d]).mean(1, keepdim=True)
        return V

    def concat_volume(self, L: torch.Tensor, R: torch.Tensor, D: int) -> torch.Tensor:
        """`build_concat_volume` (`submodule.py:129-140`): the left half is the unmasked left feature at every d."""
        B, C, H, W = L.shape
        V = L.new_zeros(B, 2 * C, D, H, W)
        for d in range(D):
            V[:, :C, d] = L
            if d < W:
                V[:, C:, d, :, d:] = R[:, :, :, : W - d]
        return V

    def substract_volume(self, L: torch.Tensor, R: torch.Tensor, D: int, G: int) -> torch.Tensor:
        """`build_substract_volume` + `groupwise_difference` (`submodule.py:104-126`)."""
        B, C, H, W = L.shape
        V = L.new_zeros(B, G, D, H, W)
        for d in range(min(D, W)):
            diff = (L[:, :, :, d:] - R[:, :, :, : W - d]).view(B, G, C // G, H, W - d)
            V[:, :, d, :, d:] = torch.pow(diff, 2).sum(2)
        return V

    def gwc_volume_norm(self, L: torch.Tensor, R: torch.Tensor, D: int, G: int) -> torch.Tensor:
        """`build_gwc_volume_norm` + `groupwise_correlation_norm` (`submodule.py:163-184`): features L2-normalised per
        group and pixel (the norm does not depend on the disparity, so it is taken once)."""
        B, C, H, W = L.shape
        Lg, Rg = L.view(B, G, C // G, H, W), R.view(B, G, C // G, H, W)
        Ln = Lg / (torch.norm(Lg, 2, 2, True) + 1e-05)
        Rn = Rg / (torch.norm(Rg, 2, 2, True) + 1e-05)
        V = L.new_zeros(B, G, D, H, W)
        for d in range(min(D, W)):
            V[:, :, d, :, d:] = (Ln[..., d:] * Rn[..., : W - d]).mean(dim=2)
        return V

    # ------------------------------------------------------------------ hot path: 3D hourglass
    def hourglass(self, x: torch.Tensor, p: str = "aggregation_out") -> torch.Tensor:
        """`aggregation.forward` (`ESMStereo.py:165-182`)."""
        c1 = self.basic(self.basic(x, p + ".conv1.0", 2, 1), p + ".conv1.1", 1, 1)
        c2 = self.basic(self.basic(c1, p + ".conv2.0", 2, 1), p + ".conv2.1", 1, 1)
        c3 = self.basic(self.basic(c2, p + ".conv3.0", 2, 1), p + ".conv3.1", 1, 1)
        u3 = self.basic(c3, p + ".conv3_up", 2, 1, deconv=True)
        u3 = u3[:, :, : c2.shape[2], : c2.shape[3], : c2.shape[4]]  # crop-to-skip, :172
        c2 = self.basic(self.basic(torch.cat((u3, c2), 1), p + ".agg_0.0", 1, 0), p + ".agg_0.1", 1, 1)
        u2 = self.basic(c2, p + ".conv2_up", 2, 1, deconv=True)
        u2 = u2[:, :, : c1.shape[2], : c1.shape[3], : c1.shape[4]]  # :177
        c1 = self.basic(self.basic(torch.cat((u2, c1), 1), p + ".agg_1.0", 1, 0), p + ".agg_1.1", 1, 1)
        return self.basic(c1, p + ".conv1_up", 2, 1, deconv=True, bn=False, act=None)

    # ------------------------------------------------------------------ hot path: regression
    @staticmethod
    def regression_top2(cost: torch.Tensor):
        """`regression_topk(cost, arange, 2)` (`submodule.py:218-225`, samples `ESMStereo.py:719-720`).
        Ties: lower disparity index wins (CPU `sort(descending)` behaviour).
        Returns (pred [B,1,H,W], idx [B,2,H,W])."""
        _, ind = torch.sort(cost, dim=1, descending=True, stable=True)
        idx = ind[:, :2]
        top = torch.gather(cost, 1, idx)
        prob = F.softmax(top, 1)
        pred = torch.sum(idx.to(cost.dtype) * prob, 1, keepdim=True)
        return pred, idx

    @staticmethod
    def disparity_regression(cost: torch.Tensor) -> torch.Tensor:
        """`disparity_regression` (`submodule.py:211-216`) -- NO softmax (callers `ESMStereo.py:725,730`)."""
        D = cost.shape[1]
        dv = torch.arange(0, D, dtype=cost.dtype, device=cost.device).view(1, D, 1, 1)
        return torch.sum(cost * dv, 1, keepdim=True)

    # ------------------------------------------------------------------ hot path: ShuffleMixer upsampler
    def layer_norm_c(self, x: torch.Tensor, wkey: str) -> torch.Tensor:
        """Bias-free LayerNorm over channels per pixel (`shufflemixer.py:40-62,83-93`)."""
        mu = x.mean(1, keepdim=True)
        var = x.var(1, keepdim=True, unbiased=False)
        return (x - mu) / torch.sqrt(var + 1e-5) * self.w(wkey).view(1, -1, 1, 1)

    def split_mlp(self, x: torch.Tensor, p: str) -> torch.Tensor:
        """`SplitPointMlp.forward` (`shufflemixer.py:33-37`): MLP on first half, channel shuffle g=8."""
        C = x.shape[1]
        a, b = x[:, : C // 2], x[:, C // 2:]
        a = self.conv(a, p + ".fc.0.weight", 1, 0, bias=p + ".fc.0.bias")
        a = self.conv(F.silu(a), p + ".fc.2.weight", 1, 0, bias=p + ".fc.2.bias")
        y = torch.cat((a, b), 1)
        B, _, H, W = y.shape
        g = 8  # 'b (g d) h w -> b (d g) h w'
        return y.view(B, g, C // g, H, W).transpose(1, 2).reshape(B, C, H, W)

    def sm_layer(self, x: torch.Tensor, p: str) -> torch.Tensor:
        """`SMLayer.forward` (`shufflemixer.py:108-112`)."""
        x = self.split_mlp(self.layer_norm_c(x, p + ".norm1.body.weight"), p + ".mlp1") + x
        k = self.w(p + ".spatial.weight").shape[-1]
        x = self.conv(x, p + ".spatial.weight", 1, k // 2, bias=p + ".spatial.bias", groups=x.shape[1])
        return self.split_mlp(self.layer_norm_c(x, p + ".norm2.body.weight"), p + ".mlp2") + x

    def fm_block(self, x: torch.Tensor, p: str) -> torch.Tensor:
        """`FMBlock.forward` (`shufflemixer.py:129-132`)."""
        x = self.sm_layer(self.sm_layer(x, p + ".net.0"), p + ".net.1") + x
        y = self.conv(x, p + ".conv.0.weight", 1, 1, bias=p + ".conv.0.bias")
        y = self.conv(F.silu(y), p + ".conv.2.weight", 1, 0, bias=p + ".conv.2.bias")
        return y + x

    def disp_mlp(self, d: torch.Tensor, p: str) -> torch.Tensor:
        """`dm2x/dm4x/dm8x/cm`: k5 p1, k3, k3, k1 p1 (`ESMStereo.py:250-253`)."""
        d = self.basic(d, p + ".0", 1, 1)
        d = self.basic(d, p + ".1", 1, 1)
        d = self.basic(d, p + ".2", 1, 1)
        return self.basic(d, p + ".3", 1, 1)

    def refine(self, disp: torch.Tensor, f1: torch.Tensor, f2: torch.Tensor, p: str) -> torch.Tensor:
        """`up_refinement.forward` (`ESMStereo.py:221-239`)."""
        c1 = self.basic(self.basic(disp, p + ".conv1.0", 2, 1), p + ".conv1.1", 1, 1)
        c2 = self.basic(self.basic(c1, p + ".conv2.0", 2, 1), p + ".conv2.1", 1, 1)
        c3 = self.basic(self.basic(c2, p + ".conv3.0", 2, 1), p + ".conv3.1", 1, 1)
        u3 = self.basic(c3, p + ".conv3_up", 2, 1, deconv=True)
        u3 = u3[:, : c2.shape[1], : c2.shape[2], : c2.shape[3]]  # :230 (only this one is cropped)
        c2 = self.basic(self.basic(torch.cat((u3, c2, f1), 1), p + ".agg_0.0", 1, 0), p + ".agg_0.1", 1, 1)
        u2 = self.basic(c2, p + ".conv2_up", 2, 1, deconv=True)
        c1 = self.basic(self.basic(torch.cat((u2, c1, f2), 1), p + ".agg_1.0", 1, 0), p + ".agg_1.1", 1, 1)
        return self.basic(c1, p + ".conv1_up", 2, 1, deconv=True, bn=False, act=None)

    def up_stage(self, prev: torch.Tensor, feat: torch.Tensor, r1: torch.Tensor, r2: torch.Tensor,
                 tag: str, with_blocks: bool, shuffle: int, p: str = "upsample_module") -> torch.Tensor:
        """One stage of `upsample4/8/16.forward` (`ESMStereo.py:296-318,396-428,484-509`):
        dm -> cat(feature) -> spx -> [to_feat -> FMBlocks] -> 1x1 -> PixelShuffle -> SiLU -> tail ->
        up_refinement -> + bilinear(prev)."""
        x = self.disp_mlp(prev, "%s.dm%s" % (p, tag))
        x = self.conv_bn_act_seq(torch.cat((x, feat), 1), "%s.spx_%s" % (p, tag), "gelu")
        if with_blocks:
            x = self.conv(x, p + ".to_feat.weight", 1, 1)
            x = self.fm_block(self.fm_block(x, p + ".blocks.0"), p + ".blocks.1")
        n = tag[0]  # '2','4','8'
        x = self.conv(x, "%s.upsampling%s.0.weight" % (p, n), 1, 0, bias="%s.upsampling%s.0.bias" % (p, n))
        x = F.silu(F.pixel_shuffle(x, shuffle))
        x = self.conv(x, "%s.tail%s.weight" % (p, tag), 1, 1, bias="%s.tail%s.bias" % (p, tag))
        x = self.refine(x, r1, r2, "%s.ref%s" % (p, tag))
        return F.interpolate(prev, scale_factor=shuffle, mode="bilinear", align_corners=False) + x

    # ------------------------------------------------------------------ hot path: confidence head
    def conf_upsample(self, feat: torch.Tensor, conf: torch.Tensor, p: str) -> torch.Tensor:
        """`conf_upsample.forward` (`ESMStereo_confidence.py:532-548`)."""
        x = self.disp_mlp(conf, p + ".cm")
        x = self.conv_bn_act_seq(torch.cat((x, feat), 1), p + ".conf_spx_4", "relu")
        x = self.conv(x, p + ".conf_spx.weight", 4, 0, deconv=True, bias=p + ".conf_spx.bias")
        sfm = F.softmax(x, 1)
        b, _, h, w = conf.shape
        nb = F.unfold(conf, 3, 1, 1).reshape(b, 9, h, w)
        nb = F.interpolate(nb, (h * 4, w * 4), mode="nearest")
        c1 = (nb * sfm).sum(1, keepdim=True)
        y = self.basic(c1, p + ".conv1", 1, 1)
        y = self.basic(y, p + ".conv2", 2, 1)
        y = self.basic(y, p + ".conv1_up", 2, 1, deconv=True)
        return y + c1

    def _cbr(self, x, p, conv, bn, pad, relu=True, stride=1):
        y = self.bn(self.conv(x, "%s.%s.weight" % (p, conv), stride, pad, bias="%s.%s.bias" % (p, conv)),
                    "%s.%s" % (p, bn))
        return F.relu(y) if relu else y

    def lafnet(self, cost, disp, imag, f1, f2, p: str = "confidence_net"):
        """`LAFNet_ESM.forward` (`ESMStereo_confidence.py:651-744`).  Returns dict of its stages."""
        out = {}
        nrm = torch.sqrt((cost ** 2).sum(1, keepdim=True) + 1e-6)  # L2normalize :645-649
        x = F.softmax(-(cost / nrm) * 100, 1)
        x = torch.topk(x, k=7, dim=1).values
        out["conf_top7"] = x
        tower = {}
        for name, src, in (("cost", x), ("disp", disp), ("imag", imag)):
            t = self._cbr(src, p, name + "_conv1", name + "_bn1", 1)
            t = self._cbr(t, p, name + "_conv2", name + "_bn2", 1)
            tower[name] = self._cbr(t, p, name + "_conv3", name + "_bn3", 0)
        att = []
        for name in ("cost", "disp", "imag"):
            t = self._cbr(tower[name], p, name + "_att_conv1", name + "_att_bn1", 1)
            att.append(self._cbr(t, p, name + "_att_conv2", name + "_att_bn2", 0, relu=False))
        att = F.softmax(torch.cat(att, 1), 1)
        x = torch.cat([tower[n] * att[:, i:i + 1] for i, n in enumerate(("cost", "disp", "imag"))], 1)
        feat = self._cbr(x, p, "embed_conv1", "embed_bn1", 1)
        out["conf_feat"] = feat
        t = self._cbr(feat, p, "scale_conv1", "scale_bn1", 1)
        t = self._cbr(t, p, "scale_conv2", "scale_bn2", 1)
        scale = 2 * torch.sigmoid(self._cbr(t, p, "scale_conv3", "scale_bn3", 0, relu=False))
        out["conf_scale"] = scale
        b, c, h, w = disp.shape
        # sampling grid :695-715 -- note y offsets are +-scale in normalised units, x offsets use
        # step_y = 2/(w-1); `step_x` is computed but unused in the reference.
        gw, gh = np.meshgrid(np.linspace(-1, 1, w), np.linspace(-1, 1, h))
        gw = torch.tensor(gw, dtype=torch.float32).to(self.dtype).to(self.device).view(1, h, w, 1).expand(b, h, w, 1)
        gh = torch.tensor(gh, dtype=torch.float32).to(self.dtype).to(self.device).view(1, h, w, 1).expand(b, h, w, 1)
        grid = torch.cat((gw, gh), 3)
        st = scale.permute(0, 2, 3, 1)
        step_y = 2 / (w - 1)
        big = torch.zeros(b, 3 * h, 3 * w, 2, dtype=self.dtype).to(self.device)
        for iy, oy in enumerate((-1, 0, 1)):
            for ix, ox in enumerate((-1, 0, 1)):
                big[:, iy::3, ix::3, :] = grid + torch.cat((ox * step_y * st, oy * st), 3)
        samp = F.grid_sample(feat, big, mode="bilinear", padding_mode="zeros", align_corners=True)
        feat = self._cbr(samp, p, "embed_conv2", "embed_bn2", 0, stride=3)
        out["conf_embed"] = feat
        o = torch.zeros(b, c, h, w, dtype=self.dtype).to(self.device) + 0.5
        for it in (1, 2, 3):  # shared convs, per-iteration BN :725-739
            t = self._cbr(torch.cat((feat, o), 1), p, "fusion_conv1", "fusion_bn1_iter%d" % it, 1)
            t = self._cbr(t, p, "fusion_conv2", "fusion_bn2_iter%d" % it, 1)
            o = self._cbr(t, p, "fusion_conv3", "fusion_bn3_iter%d" % it, 0)
        out["conf_init"] = o
        o4 = self.conf_upsample(f1, o, p + ".conf_up4")
        out["conf_4"] = o4
        o1 = self.conf_upsample(f2, o4, p + ".conf_up1")
        out["conf"] = torch.sigmoid(o1)
        return out

    # ------------------------------------------------------------------ hot path, from descriptors on
    def hot_path(self, mL, mR, featsL, stemsL) -> Dict[str, torch.Tensor]:
        """`ESMStereo.forward` from the volume on (`ESMStereo.py:700-745`; `_confidence.py:938-974`)."""
        s, D, out = self.s, self.maxdisp // self.s, {}
        att = None
        if s == 16:  # `semantic` :606-618, :697
            att = self.conv(self.basic(featsL[3], "semantic.0", 1, 1), "semantic.1.weight", 1, 1).unsqueeze(2)
            out["att"] = att
        if self.ncorr:
            vol = self.norm_corr_volume(mL, mR, D)
            out["volume"] = vol
            vol = self.basic(vol, "corr_stem", 1, 1)
            out["stem"] = vol  # module output; for cv16 the `* att` comes after (`ESMStereo.py:703`)
            if s == 16:
                vol = vol * att
        if self.gwc:
            vol = self.gwc_volume(mL, mR, D, 32)
            out["volume"] = vol
            vol = self.basic(vol * att if s == 16 else vol, "group_stem", 1, 1)
            out["stem"] = vol
        vol = self.basic(vol, "agg", 1, 1)
        out["agg"] = vol
        cost = self.hourglass(vol).squeeze(1)
        out["cost"] = cost
        if s == 4:
            init, idx = self.regression_top2(cost)
            out["top2_idx"] = idx
            d2 = self.up_stage(init, featsL[0], featsL[1], featsL[0], "2x", True, 2)
            d1 = self.up_stage(d2, stemsL[0], featsL[0], stemsL[0], "4x", False, 2)
            scales = [d1, d2]
        elif s == 8:
            init = self.disparity_regression(cost)
            d4 = self.up_stage(init, featsL[1], featsL[2], featsL[1], "2x", True, 2)
            d2 = self.up_stage(d4, featsL[0], featsL[1], featsL[0], "4x", False, 2)
            d1 = self.up_stage(d2, stemsL[0], featsL[0], stemsL[0], "8x", False, 2)
            scales = [d1, d2, d4]
        else:
            init = self.disparity_regression(cost)
            f2 = self.basic(featsL[3], "conv_f2", 1, 1)
            f0 = self.basic(featsL[0], "conv_f0", 1, 1)
            d2 = self.up_stage(init, f2, f2, featsL[2], "2x", True, 4)
            d1 = self.up_stage(d2, featsL[1], featsL[1], f0, "4x", False, 4)
            scales = [d1, d2]
        out["init_pred"] = init
        out["scales"] = [t.squeeze(1) * 4 for t in scales]  # every scale is *4, :737-745
        out["disp"] = out["scales"][0]
        if self.confidence and s == 16:
            out.update(self.lafnet(cost, init, mL, featsL[3], featsL[1]))
            out["conf"] = out["conf"].squeeze(1)
        return out

    def forward(self, left: torch.Tensor, right: torch.Tensor) -> Dict[str, torch.Tensor]:
        left, right = left.to(self.dtype), right.to(self.dtype)
        with torch.no_grad():
            fL, sL, mL = self.descriptors(left)
            _, _, mR = self.descriptors(right)
            out = self.hot_path(mL, mR, fL, sL)
        out["match_left"], out["match_right"] = mL, mR
        out["feats_left"], out["stems_left"] = fL, sL
        return out

    __call__ = forward

    def calibrate(self, left: torch.Tensor, right: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One train-mode-BN pass that overwrites every running_mean/var with batch statistics
        (the reference model in `.train()` with every BN `momentum=1.0`), which makes the
        random-weight cost volume well conditioned.  Shared 2D layers see the
        left image first and the right image second (`ESMStereo.py:640-659`), so -- as in the
        reference -- the right image's statistics are the ones that survive in those buffers; the
        normalisation of each batch depends only on that batch, so whole-branch order is equivalent
        to the reference's per-block interleaving.  Returns the updated state_dict."""
        self.calibrating = True
        try:
            with torch.no_grad():
                left, right = left.to(self.dtype), right.to(self.dtype)
                fL, sL, mL = self.descriptors(left)
                _, _, mR = self.descriptors(right)
                self.hot_path(mL, mR, fL, sL)
        finally:
            self.calibrating = False
        return self.sd


# ----------------------------------------------------------------------------------------------------
# pre / post-processing of the reference's scripts, restated on torch-CPU tensors
# ----------------------------------------------------------------------------------------------------
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def preprocess_images(rgb_u8: torch.Tensor, size, mode: str = "test_kitti") -> torch.Tensor:
    """uint8 [B,h,w,3] -> float32 [B,3,Hp,Wp].  "test_kitti": `limg.crop((w - wi, h - hi, w, h))` then ToTensor +
    Normalize (`test_kitti.py:93-106`; PIL fills the area outside the image with black, which is then normalised).
    "kitti_dataset": ToTensor + Normalize, then `np.lib.pad(((0,0),(top_pad,0),(0,right_pad)), constant 0)`
    (`datasets/kitti_dataset.py:145-160`)."""
    B, h, w, _ = rgb_u8.shape
    Hp, Wp = size
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(1, 3, 1, 1)
    if mode == "test_kitti":
        canvas = torch.zeros(B, Hp, Wp, 3, dtype=torch.uint8)
        canvas[:, Hp - h:, Wp - w:] = rgb_u8
        t = canvas.permute(0, 3, 1, 2).to(torch.float32).div(255)  # ToTensor
        return t.sub(mean).div(std)                                # Normalize
    t = rgb_u8.permute(0, 3, 1, 2).to(torch.float32).div(255).sub(mean).div(std)
    out = torch.zeros(B, 3, Hp, Wp, dtype=torch.float32)
    out[:, :, Hp - h:, :w] = t
    return out


def disparity_to_uint16(disp: torch.Tensor, size, mode: str = "test_kitti") -> torch.Tensor:
    """`pred_disp[:, hi - h:, wi - w:]` (`test_kitti.py:114`) or `disp_est[top_pad:, :-right_pad]` (`save_disp.py:83`),
    then `np.round(d * 256).astype(np.uint16)` (`test_kitti.py:127`, `save_disp.py:87`)."""
    import numpy as np
    B, Hp, Wp = disp.shape
    h, w = size
    crop = disp[:, Hp - h:, Wp - w:] if mode == "test_kitti" else disp[:, Hp - h:, :w]
    return torch.from_numpy(np.round(crop.numpy() * 256).astype(np.uint16).astype(np.int32))
