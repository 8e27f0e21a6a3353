"""The oracle (oracle/esm_oracle.py) is held to vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle.esm_oracle import EsmOracle
from tests.helpers import (GOLDEN_NAMES, golden_blob, golden_config, golden_inputs, golden_state_dict,
                           rel_err, sample)


def _oracle(name, sd=None, dtype=torch.float32):
    cfg = golden_config(name)
    sd = golden_state_dict(name) if sd is None else sd
    return EsmOracle(sd, 192, cfg["gwc"], cfg["norm_correlation"], cfg["backbone"], cfg["cv_scale"],
                     confidence=cfg["model"] == "ESMStereo_confidence", dtype=dtype)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_reference_stages(name):
    blob = golden_blob(name)
    left, right = golden_inputs(name)
    out = _oracle(name)(left, right)
    # same torch CPU kernels in the same order -> expect (near) bit equality; allow fp32 noise
    assert rel_err(sample(out["match_left"]), blob["match_left_sample"]) < 1e-5
    assert rel_err(sample(out["volume"]), blob["volume_sample"]) < 1e-5
    assert rel_err(out["volume"].sum((3, 4)).numpy(), blob["volume_sum_over_hw"]) < 1e-5
    assert rel_err(sample(out["stem"]), blob["stem_sample"]) < 1e-5
    assert rel_err(sample(out["agg"]), blob["agg_sample"]) < 1e-5
    assert rel_err(out["cost"].numpy(), blob["cost"]) < 1e-5
    assert rel_err(out["init_pred"].numpy(), blob["init_pred"]) < 1e-5
    if "top2_idx" in blob.files:
        assert np.array_equal(out["top2_idx"].numpy().astype(np.int16), blob["top2_idx"])
    assert rel_err(out["disp"].numpy(), blob["disp"]) < 1e-5
    epe = float(np.abs(out["disp"].numpy() - blob["disp"]).mean())
    assert epe < 1e-3, epe
    i = 0
    while "train_out_%d" % i in blob.files:
        assert rel_err(out["scales"][i].numpy(), blob["train_out_%d" % i]) < 1e-5
        i += 1
    if "conf" in blob.files:
        assert rel_err(out["conf_embed"].numpy(), np.maximum(blob["conf_embed2_bn"], 0)) < 1e-5
        assert rel_err(out["conf_init"].numpy(), blob["conf_init"]) < 1e-5
        assert rel_err(out["conf_4"].numpy(), blob["conf_4"]) < 1e-5
        assert float(np.abs(out["conf"].numpy() - blob["conf"]).max()) < 1e-5


@pytest.mark.parametrize("name", ["cv4_gwc", "conf16_gwc", "cv8_gwc"])
def test_oracle_calibration_matches_reference(name):
    """oracle.calibrate() reproduces the BN buffers the reference leaves after its train-mode pass."""
    blob = golden_blob(name)
    sd = golden_state_dict(name, calibrated=False)
    orc = _oracle(name, sd)
    left, right = golden_inputs(name)
    new_sd = orc.calibrate(left, right)
    worst = 0.0
    for k in blob.files:
        if k.startswith("bn/"):
            worst = max(worst, rel_err(new_sd[k[3:]].numpy(), blob[k]))
    assert worst < 1e-3, worst  # batch-stat reductions differ in summation order from nn.BatchNorm


def test_oracle_fp64_agrees_with_fp32():
    left, right = golden_inputs("cv4_gwc")
    o32 = _oracle("cv4_gwc")(left, right)
    o64 = _oracle("cv4_gwc", dtype=torch.float64)(left, right)
    assert float((o32["disp"].double() - o64["disp"]).abs().mean()) < 1e-3
    mism = (o32["top2_idx"].sort(1).values != o64["top2_idx"].sort(1).values).float().mean().item()
    assert mism < 0.002, mism


def test_regression_tie_break_lower_index_wins():
    cost = torch.zeros(1, 6, 1, 2)
    cost[0, 4, 0, 0] = 1.0  # pixel 0: unique max at 4, tie for second among {0,1,2,3,5} -> 0
    pred, idx = EsmOracle.regression_top2(cost)
    assert idx[0, :, 0, 0].tolist() == [4, 0]
    assert idx[0, :, 0, 1].tolist() == [0, 1]  # all equal -> 0 then 1
