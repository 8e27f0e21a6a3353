"""Backbones behind `Feature` (`/root/reference/models/ESMStereo.py:40-77`), which the reference pulls from timm
(`timm.create_model('efficientnet_b2' | 'mobilenetv2_100', pretrained=True, features_only=True)`, :46,55).

`make_backbone(name)` returns, in this order (ESM_BACKBONE = auto | timm | compat | standin overrides):
  1. a real `timm` model when timm is importable ("torch" engine: its own cuDNN forward);
  2. otherwise `timm_compat.TimmCompatBackbone`: the same two architectures restated with timm's module / parameter
     names, so reference checkpoints load by key, running on libesm_b200 kernels;
  3. the structural STAND-IN below only on explicit request (ESM_BACKBONE=standin).  It has the attribute surface the
     reference touches (`conv_stem`, `bn1`, `blocks[0:7]`) and the stage widths / strides of the real backbones but
     plain conv-bn-relu6 stages, i.e. DIFFERENT parameter names: a real checkpoint loaded through the reference's
     key filter (test_kitti.py:57-61) would silently leave it at random weights.  It exists because the reference
     goldens (tests/golden/timm_shim), the CPU oracle and bench.py's synthetic workload were defined on it, with
     timm absent from the image; those callers opt in explicitly and a warning says so.
"""
import os
import warnings

import torch.nn as nn

# (out_channels, stride) of timm's 7 stages; reference slices blocks[0:1],[1:2],[2:3],[3:5],[5:6]
STAGES = {
    "efficientnet_b2": [(16, 1), (24, 2), (48, 2), (88, 2), (120, 1), (208, 2), (352, 1)],
    "mobilenetv2_100": [(16, 1), (24, 2), (32, 2), (64, 2), (96, 1), (160, 2), (320, 1)],
}
# channels the reference reads back (`self.chans`, ESMStereo.py:48,57)
FEATURE_CHANS = {
    "efficientnet_b2": [16, 24, 48, 120, 208],
    "mobilenetv2_100": [16, 24, 32, 96, 160],
}


class StandInBackbone(nn.Module):
    """conv_stem (3->32, k3 s2) + bn1 + 7 conv-bn-relu6 stages with the real stage widths."""

    def __init__(self, name: str) -> None:
        super().__init__()
        if name not in STAGES:
            raise ValueError("unknown backbone %r" % (name,))
        self.conv_stem = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        stages, cin = [], 32
        for cout, stride in STAGES[name]:
            stages.append(nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1, bias=False),
                                        nn.BatchNorm2d(cout), nn.ReLU6()))
            cin = cout
        self.blocks = nn.Sequential(*stages)


def create_model(name, pretrained=False, features_only=True, **_unused):
    """Signature-compatible with `timm.create_model` as called by the reference (the stand-in)."""
    return StandInBackbone(name)


def make_backbone(name: str) -> nn.Module:
    mode = os.environ.get("ESM_BACKBONE", "auto").lower()
    if mode not in ("auto", "timm", "compat", "standin"):
        raise ValueError("ESM_BACKBONE must be auto, timm, compat or standin (got %r)" % mode)
    if mode == "standin":
        warnings.warn("esmstereo_b200: ESM_BACKBONE=standin -- `feature.*` is a structural stand-in whose parameter names differ from "
                      "timm's; real ESMStereo checkpoints will NOT fill it (use the default backbone for them)", stacklevel=2)
        return StandInBackbone(name)
    if mode in ("auto", "timm"):
        try:
            import timm  # type: ignore
            if getattr(timm, "__esm_b200_shim__", False):
                raise ImportError("timm shim")
            return timm.create_model(name, pretrained=False, features_only=True)
        except ImportError:
            if mode == "timm":
                raise
    from .timm_compat import TimmCompatBackbone
    return TimmCompatBackbone(name)
