"""timm-key-compatible backbones (ESMStereo.py:40-77): parameter names / shapes as the
reference's checkpoints carry them, the loud checkpoint loader, and -- on the GPU -- the libesm_b200 execution of the
blocks (1x1 on the conv engines, depthwise / pooling / SE kernels) against the modules' own PyTorch forward."""
import contextlib
import io
import os

import pytest
import torch


@contextlib.contextmanager
def _backbone(mode):
    old = os.environ.get("ESM_BACKBONE")
    os.environ["ESM_BACKBONE"] = mode
    try:
        yield
    finally:
        if old is None:
            os.environ.pop("ESM_BACKBONE", None)
        else:
            os.environ["ESM_BACKBONE"] = old


def _model(name, backbone, cv, mode="compat"):
    from esmstereo_b200 import __models__
    with _backbone(mode), contextlib.redirect_stdout(io.StringIO()):
        return __models__[name](192, True, False, backbone, cv)


def test_efficientnet_b2_keys_follow_timm():
    from esmstereo_b200.timm_compat import TimmCompatBackbone, stage_channels
    m = TimmCompatBackbone("efficientnet_b2")
    sd = m.state_dict()
    assert stage_channels("efficientnet_b2") == [16, 24, 48, 88, 120, 208, 352]
    assert [len(s) for s in m.blocks] == [2, 3, 3, 4, 4, 5, 2]  # B0 repeats x depth multiplier 1.2, rounded up: 23 blocks
    want = {
        "conv_stem.weight": (32, 3, 3, 3), "bn1.running_var": (32,),
        # stage 0: DepthwiseSeparableConv, SE reduced to a quarter of the block input
        "blocks.0.0.conv_dw.weight": (32, 1, 3, 3), "blocks.0.0.se.conv_reduce.weight": (8, 32, 1, 1), "blocks.0.0.se.conv_reduce.bias": (8,),
        "blocks.0.0.se.conv_expand.weight": (32, 8, 1, 1), "blocks.0.0.conv_pw.weight": (16, 32, 1, 1), "blocks.0.0.bn2.weight": (16,),
        "blocks.0.1.se.conv_reduce.weight": (4, 16, 1, 1),
        # stage 1: InvertedResidual, expansion 6
        "blocks.1.0.conv_pw.weight": (96, 16, 1, 1), "blocks.1.0.conv_dw.weight": (96, 1, 3, 3), "blocks.1.0.se.conv_reduce.weight": (4, 96, 1, 1),
        "blocks.1.0.conv_pwl.weight": (24, 96, 1, 1), "blocks.1.0.bn3.bias": (24,), "blocks.1.1.se.conv_reduce.weight": (6, 144, 1, 1),
        "blocks.2.0.conv_dw.weight": (144, 1, 5, 5), "blocks.2.0.conv_pwl.weight": (48, 144, 1, 1),
        "blocks.4.0.conv_pw.weight": (528, 88, 1, 1), "blocks.4.3.conv_pwl.weight": (120, 720, 1, 1),
        "blocks.5.0.conv_dw.weight": (720, 1, 5, 5), "blocks.5.4.conv_pwl.weight": (208, 1248, 1, 1),
        "blocks.6.1.conv_pwl.weight": (352, 2112, 1, 1),
    }
    for k, shape in want.items():
        assert k in sd and tuple(sd[k].shape) == shape, k
    # nothing but timm's leaf names
    leaves = {k.split(".")[-2] for k in sd if k.startswith("blocks.")}
    assert leaves <= {"conv_dw", "conv_pw", "conv_pwl", "bn1", "bn2", "bn3", "conv_reduce", "conv_expand"}


def test_mobilenetv2_keys_follow_timm():
    from esmstereo_b200.timm_compat import TimmCompatBackbone
    m = TimmCompatBackbone("mobilenetv2_100")
    sd = m.state_dict()
    assert [len(s) for s in m.blocks] == [1, 2, 3, 4, 3, 3, 1]
    assert not any(".se." in k for k in sd)
    assert tuple(sd["blocks.0.0.conv_pw.weight"].shape) == (16, 32, 1, 1)
    assert tuple(sd["blocks.1.0.conv_pw.weight"].shape) == (96, 16, 1, 1)
    assert tuple(sd["blocks.4.2.conv_pwl.weight"].shape) == (96, 576, 1, 1)
    assert tuple(sd["blocks.5.0.conv_dw.weight"].shape) == (576, 1, 3, 3)


def test_model_keys_are_the_reference_layout_and_loader_is_loud():
    """`feature.block3` = timm stages 3 and 4 (ESMStereo.py:62-66); a checkpoint that does not fill the backbone raises."""
    from esmstereo_b200 import load_reference_checkpoint
    m = _model("ESMStereo", "efficientnet_b2", 4)
    keys = set(m.state_dict())
    for k in ("feature.conv_stem.weight", "feature.bn1.running_mean", "feature.block0.0.0.conv_dw.weight",
              "feature.block3.0.0.conv_pw.weight", "feature.block3.1.3.bn3.weight", "feature.block4.0.4.conv_pwl.weight",
              "group_stem.conv.weight", "upsample_module.blocks.0.net.0.mlp1.fc.0.weight"):
        assert k in keys, k
    assert not any(k.startswith("feature.block4.1") for k in keys)  # stage 6 is never kept (blocks[0:6])
    assert m.feature.chans == [16, 24, 48, 120, 208] and m.feature_engine == "esm"
    # a DataParallel checkpoint of the same architecture loads, prefix or not
    ckpt = {"module." + k: torch.full_like(v, 0.25) if v.is_floating_point() else v for k, v in m.state_dict().items()}
    assert load_reference_checkpoint(m, ckpt) == []
    assert float(m.feature.conv_stem.weight.detach().flatten()[0]) == 0.25
    # a checkpoint whose backbone names differ (here: the stand-in's) must not load silently
    with pytest.warns(UserWarning):
        standin = _model("ESMStereo", "efficientnet_b2", 4, mode="standin")
    with pytest.raises(RuntimeError, match="backbone tensors unfilled"):
        load_reference_checkpoint(m, standin.state_dict())
    assert len(load_reference_checkpoint(m, standin.state_dict(), require_backbone=False)) > 100


def test_compat_torch_forward_shapes():
    from esmstereo_b200.feature2d import Feature
    with _backbone("compat"):
        for name, chans in (("efficientnet_b2", [16, 24, 48, 120, 208]), ("mobilenetv2_100", [16, 24, 32, 96, 160])):
            f = Feature(name).eval()
            with torch.no_grad():
                outs = f(torch.randn(1, 3, 64, 96), "torch")
            assert [o.shape[1] for o in outs] == chans
            assert [o.shape[-1] for o in outs] == [48, 24, 12, 6, 3]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["efficientnet_b2", "mobilenetv2_100"])
def test_compat_esm_engine_matches_torch_engine(name):
    """Depthwise + SE + 1x1 on libesm_b200 vs the same modules through cuDNN (TF32 off), random weights."""
    from esmstereo_b200.feature2d import Feature
    from esmstereo_b200.weights import fill_deterministic
    with _backbone("compat"):
        f = Feature(name)
    f.load_state_dict(fill_deterministic(f.state_dict(), seed=3))
    f = f.cuda().eval()
    x = torch.randn(2, 3, 96, 160, generator=torch.Generator().manual_seed(0)).cuda()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            want = f(x, "torch")
            got = f(x, "esm")
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-4 * float(b.abs().max()) + 1e-6


@pytest.mark.gpu
def test_model_with_compat_backbone_runs_and_matches_torch_feature_side():
    m = _model("ESMStereo", "efficientnet_b2", 4)
    from esmstereo_b200.weights import fill_deterministic, synthetic_pair
    m.load_state_dict(fill_deterministic(m.state_dict(), seed=1))
    m = m.cuda().eval()
    l, r = [t.cuda() for t in synthetic_pair(1, 128, 256, shift=7, seed=0)]
    m.capture = {}
    d_esm = m(l, r, train_status=False)[-1]
    ml_esm = m.capture["match_left"].clone()
    m.feature_engine = "torch"
    m.capture = {}
    d_torch = m(l, r, train_status=False)[-1]
    ml_torch = m.capture["match_left"]
    assert d_esm.shape == d_torch.shape == (1, 128, 256) and torch.isfinite(d_esm).all()
    # the matching descriptors (backbone + FeatUp + stems + desc) from the two engines of the 2D side; the disparities of
    # an uncalibrated random network are not compared (near-tie top-2 flips amplify 1e-6 differences)
    assert float((ml_esm - ml_torch).abs().max()) <= 1e-3 * float(ml_torch.abs().max())
