"""BASELINE.json shapes against the REAL reference: tests/golden/make_golden.py ran /root/reference's ESMStereo at
384x1248 (config B) and ESMStereo_confidence at 992x1472 (config E) on CPU and stored samples of the stage outputs
(`kitti_cv4_gwc.npz`, `conf16_gwc_full.npz`).  The oracle is held to them on CPU, the CUDA path on the GPU -- so the
full-size parity no longer rests on the port alone."""
import contextlib
import io

import numpy as np
import pytest
import torch

from esmstereo_b200.weights import synthetic_pair
from oracle.esm_oracle import EsmOracle

from .helpers import golden_blob, golden_config, golden_state_dict, rel_err, sample

FULL = ["kitti_cv4_gwc", "conf16_gwc_full"]


def _inputs(cfg):
    return synthetic_pair(1, cfg["H"], cfg["W"], shift=23, seed=0)


@pytest.mark.parametrize("name", FULL)
def test_oracle_matches_reference_at_full_size(name):
    cfg, blob = golden_config(name), golden_blob(name)
    conf = cfg["model"] == "ESMStereo_confidence"
    orc = EsmOracle(golden_state_dict(name), 192, cfg["gwc"], cfg["norm_correlation"], cfg["backbone"], cfg["cv_scale"], confidence=conf)
    out = orc(*_inputs(cfg))
    assert rel_err(sample(out["match_left"]), blob["match_left_sample"]) < 1e-5
    assert rel_err(sample(out["stem"]), blob["stem_sample"]) < 1e-5
    assert rel_err(sample(out["cost"].unsqueeze(1)), blob["cost_sample"]) < 2e-5
    if "top2_idx" in blob.files:
        a = np.sort(out["top2_idx"].numpy(), 1)
        b = np.sort(blob["top2_idx"].astype(np.int64), 1)
        assert (a != b).mean() <= 1e-4  # CPU fp32 vs CPU fp32, different summation order in the port: near ties only
    assert float(np.abs(out["disp"][:, ::4, ::4].numpy() - blob["disp_q"]).mean()) <= 0.01
    if conf:
        assert float(np.abs(out["conf"][:, ::4, ::4].numpy() - blob["conf_q"]).mean()) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", FULL)
def test_gpu_matches_reference_at_full_size(name):
    from esmstereo_b200 import __models__
    cfg, blob = golden_config(name), golden_blob(name)
    conf = cfg["model"] == "ESMStereo_confidence"
    with contextlib.redirect_stdout(io.StringIO()):
        args = (192, cfg["gwc"], cfg["norm_correlation"], cfg["backbone"], cfg["cv_scale"])
        m = __models__[cfg["model"]](*args)
    m.load_state_dict(golden_state_dict(name))
    m = m.cuda().eval()
    m.capture = {}
    left, right = [t.cuda() for t in _inputs(cfg)]
    if conf:
        disp, cf = m(left, right)
    else:
        disp, cf = m(left, right, train_status=False)[-1], None
    cap = m.capture
    assert rel_err(sample(cap["match_left"]), blob["match_left_sample"]) < 1e-4
    assert rel_err(sample(cap["stem"]), blob["stem_sample"]) < 1e-4
    # cost volume within 1e-4 relative (BASELINE.json north_star), against the reference itself
    assert rel_err(sample(cap["cost"].unsqueeze(1)), blob["cost_sample"]) < 1e-4
    flips = 0.0
    if "top2_idx" in blob.files:
        a = np.sort(cap["top2_idx"].cpu().numpy(), 1)
        b = np.sort(blob["top2_idx"].astype(np.int32), 1)
        flips = float((a != b).any(1).mean())
        assert flips <= 2e-4, "top-2 index mismatch fraction %g" % flips
    epe = float(np.abs(disp[:, ::4, ::4].cpu().numpy() - blob["disp_q"]).mean())
    assert epe <= (0.01 if flips == 0 else 0.1), "EPE vs the reference = %g px (flip fraction %g)" % (epe, flips)
    if conf:
        err = np.abs(cf[:, ::4, ::4].cpu().numpy() - blob["conf_q"])
        assert float(err.mean()) < 2e-3 and float(err.max()) < 0.15  # ill-conditioned head: softmax(-100 cost / |cost|)
