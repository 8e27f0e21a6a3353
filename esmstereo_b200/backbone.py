"""Stand-in for the timm backbones the reference pulls in (`/root/reference/models/ESMStereo.py:46,55`).

timm is not installed in this image and there is no network, so neither the reference nor this
package can build `timm.create_model('efficientnet_b2' | 'mobilenetv2_100', features_only=True)`.
This module provides a structural stand-in with the attribute surface the reference touches
(`conv_stem`, `bn1`, `blocks[0:7]`) and the stage channels / strides of the two real backbones,
so that `Feature` (ESMStereo.py:40-77) slices it exactly like a timm model.  It is OUTSIDE the
hot path (2D feature side stays PyTorch).  If a real `timm` is importable it
is used instead, keeping checkpoint key compatibility.
"""
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
    """Signature-compatible with `timm.create_model` as called by the reference."""
    return StandInBackbone(name)


def make_backbone(name: str) -> nn.Module:
    try:  # a real timm, when present, keeps checkpoint keys loadable
        import timm  # type: ignore
        if getattr(timm, "__esm_b200_shim__", False):
            raise ImportError
        return timm.create_model(name, pretrained=False, features_only=True)
    except ImportError:
        return StandInBackbone(name)
