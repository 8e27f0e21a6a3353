"""End-to-end GPU parity through the reference-facing API.

* golden: outputs of the REAL reference (tests/golden/*.npz, made by tests/golden/make_golden.py);
* oracle: oracle/esm_oracle.py on CPU at BASELINE.json's full shapes (KITTI 384x1248, confidence
  992x1472), with BN statistics calibrated by the oracle (validated against the reference in
  tests/test_oracle_golden.py).

Gates (BASELINE.json north_star, fp32 mode): final disparity EPE delta <= 0.01 px; top-2 indices equal
(a vanishing fraction of near-tie pixels may swap: bounded at 2e-4 and reported); stage tensors within
1e-4 relative.
"""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle.esm_oracle import EsmOracle
from tests.helpers import (GOLDEN_NAMES, golden_blob, golden_config, golden_inputs, golden_state_dict, rel_err,
                           sample)
from esmstereo_b200.weights import fill_deterministic, synthetic_pair

pytestmark = pytest.mark.gpu


def build(model_name, gwc, ncorr, backbone, s, sd=None):
    from esmstereo_b200 import __models__
    with contextlib.redirect_stdout(io.StringIO()):
        m = __models__[model_name](192, gwc, ncorr, backbone, s)
    if sd is not None:
        m.load_state_dict(sd)
    return m.cuda().eval()


def run(m, model_name, left, right, train_status=False):
    if model_name == "ESMStereo":
        return m(left, right, train_status), None
    if model_name == "ESMStereo_trt":
        return [m(left, right)], None
    d, c = m(left, right)
    return [d], c


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_matches_reference_golden(name):
    cfg, blob = golden_config(name), golden_blob(name)
    m = build(cfg["model"], cfg["gwc"], cfg["norm_correlation"], cfg["backbone"], cfg["cv_scale"], golden_state_dict(name))
    m.capture = {}
    left, right = golden_inputs(name)
    outs, conf = run(m, cfg["model"], left.cuda(), right.cuda())
    cap = m.capture
    assert rel_err(sample(cap["match_left"]), blob["match_left_sample"]) < 1e-4
    if not (cfg["norm_correlation"] and cfg["cv_scale"] == 16):  # there `* att` is fused into the stem kernel
        assert rel_err(sample(cap["stem"]), blob["stem_sample"]) < 1e-4
    assert rel_err(sample(cap["agg"]), blob["agg_sample"]) < 1e-4
    assert rel_err(cap["cost"].cpu().numpy(), blob["cost"]) < 1e-4
    if "top2_idx" in blob.files:
        mism = float((np.sort(cap["top2_idx"].cpu().numpy(), 1) != np.sort(blob["top2_idx"].astype(np.int32), 1)).mean())
        assert mism <= 2e-4, "top-2 index mismatch fraction %g" % mism
    disp = outs[0].cpu().numpy()
    assert disp.shape == blob["disp"].shape
    epe = float(np.abs(disp - blob["disp"]).mean())
    assert epe <= 0.01, "EPE delta vs reference = %g px" % epe
    if conf is not None:
        assert float(np.abs(conf.cpu().numpy() - blob["conf"]).max()) < 1e-3
    if cfg["model"] == "ESMStereo":
        outs, _ = run(m, cfg["model"], left.cuda(), right.cuda(), train_status=True)
        i = 0
        while "train_out_%d" % i in blob.files:
            assert float(np.abs(outs[i].cpu().numpy() - blob["train_out_%d" % i]).mean()) <= 0.01
            i += 1
        assert i == len(outs)


def _full_size(model_name, gwc, backbone, s, B, H, W, seed=0):
    from esmstereo_b200 import __models__
    with contextlib.redirect_stdout(io.StringIO()):
        m = __models__[model_name](192, gwc, not gwc, backbone, s)
    sd = fill_deterministic(m.state_dict(), seed=seed)
    left, right = synthetic_pair(B, H, W, shift=23, seed=seed)
    orc = EsmOracle(sd, 192, gwc, not gwc, backbone, s, confidence=model_name == "ESMStereo_confidence")
    sd = orc.calibrate(left[:1], right[:1])
    want = orc(left, right)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.capture = {}
    outs, conf = run(m, model_name, left.cuda(), right.cuda())
    return m, want, outs, conf


def test_kitti_shape_vs_oracle():
    """BASELINE.json configs[1]: 384x1248, batch 1, cv4 gwc."""
    m, want, outs, _ = _full_size("ESMStereo", True, "efficientnet_b2", 4, 1, 384, 1248)
    cap = m.capture
    assert rel_err(cap["cost"].cpu().numpy(), want["cost"].numpy()) < 1e-4
    mism = float((cap["top2_idx"].cpu().long().sort(1).values != want["top2_idx"].sort(1).values).float().mean())
    assert mism <= 2e-4, mism
    epe = float((outs[0].cpu() - want["disp"]).abs().mean())
    assert epe <= 0.01, epe
    # unfused volume path gives the same answer as the fused one
    m.fuse_volume = False
    d2 = m(*[t.cuda() for t in synthetic_pair(1, 384, 1248, shift=23, seed=0)], False)[-1]
    assert float((d2 - outs[0]).abs().max()) < 1e-3


def test_sceneflow_shape_batch_vs_oracle():
    """BASELINE.json configs[2] shape (544x960), batch 2 here to bound the CPU oracle's time."""
    m, want, outs, _ = _full_size("ESMStereo", True, "efficientnet_b2", 4, 2, 544, 960, seed=1)
    epe = float((outs[0].cpu() - want["disp"]).abs().mean())
    assert epe <= 0.01, epe


def test_confidence_highres_vs_oracle():
    """BASELINE.json configs[4]: ESMStereo_confidence, 992x1472, cv16."""
    m, want, outs, conf = _full_size("ESMStereo_confidence", True, "mobilenetv2_100", 16, 1, 992, 1472, seed=2)
    epe = float((outs[0].cpu() - want["disp"]).abs().mean())
    assert epe <= 0.01, epe
    assert float((conf.cpu() - want["conf"]).abs().max()) < 1e-3


def test_api_contract_and_graph_replay():
    """trt variant returns a bare tensor; DataParallel + 'module.' checkpoints load by key filtering
    (test_kitti.py:52-61); CUDA-graph replay equals eager."""
    from esmstereo_b200 import GraphedStereo
    name = "cv4_gwc"
    cfg, blob = golden_config(name), golden_blob(name)
    sd = golden_state_dict(name)
    left, right = [t.cuda() for t in golden_inputs(name)]
    trt = build("ESMStereo_trt", True, False, "efficientnet_b2", 4, sd)
    out = trt(left, right)
    assert isinstance(out, torch.Tensor) and out.shape == (1, cfg["H"], cfg["W"])
    assert float(np.abs(out.cpu().numpy() - blob["disp"]).mean()) <= 0.01
    with contextlib.redirect_stdout(io.StringIO()):
        from esmstereo_b200 import __models__
        dp = torch.nn.DataParallel(__models__["ESMStereo"](192, True, False, "efficientnet_b2", 4))
    dp.cuda().eval()
    ckpt = {"module." + k: v for k, v in sd.items()}
    model_dict = dp.state_dict()
    pre = {k: v for k, v in ckpt.items() if k in model_dict}
    assert len(pre) == len(model_dict)
    model_dict.update(pre)
    dp.load_state_dict(model_dict)
    eager = dp(left, right, train_status=False)[-1]
    assert float(np.abs(eager.cpu().numpy() - blob["disp"]).mean()) <= 0.01
    g = GraphedStereo(dp.module, tuple(left.shape), train_status=False)
    for _ in range(2):
        replay = g(left, right)[-1]
    torch.cuda.synchronize()
    assert torch.equal(replay, eager)


def test_cpu_inputs_fail_loudly():
    m = build("ESMStereo", True, False, "efficientnet_b2", 4)
    with pytest.raises(RuntimeError):
        m.cpu()(torch.zeros(1, 3, 64, 128), torch.zeros(1, 3, 64, 128), False)
