"""Shared test helpers: rebuild golden state_dicts from names alone, tolerance helpers."""
import json
import os

import numpy as np
import torch

from esmstereo_b200.weights import fill_deterministic, synthetic_pair

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = ["cv4_gwc", "cv4_gwc_wide", "cv4_ncorr", "cv8_gwc", "cv16_gwc", "cv16_ncorr", "conf16_gwc"]


def golden_config(name):
    with open(os.path.join(GOLDEN, name + ".keys.json")) as f:
        meta = json.load(f)
    model, gwc, ncorr, backbone, s, H, W = meta["config"]
    return dict(model=model, gwc=gwc, norm_correlation=ncorr, backbone=backbone, cv_scale=s, H=H, W=W,
                keys=meta["state_dict"])


def golden_blob(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_state_dict(name, calibrated=True):
    """Name-keyed deterministic weights + (optionally) the reference-calibrated BN buffers."""
    cfg = golden_config(name)
    sd = {}
    for k, shape in cfg["keys"].items():
        dt = torch.long if k.endswith("num_batches_tracked") else torch.float32
        sd[k] = torch.zeros(shape, dtype=dt)
    fill_deterministic(sd, seed=0)
    if calibrated:
        blob = golden_blob(name)
        for k in blob.files:
            if k.startswith("bn/"):
                sd[k[3:]] = torch.from_numpy(blob[k]).clone()
    return sd


def golden_inputs(name):
    cfg = golden_config(name)
    return synthetic_pair(1, cfg["H"], cfg["W"], shift=7, seed=0)


def sample(t, limit=20000):
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // limit)
    return flat[::step].to(torch.float32).cpu().numpy()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
