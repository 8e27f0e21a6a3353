"""Generate golden vectors by running the REAL reference (read-only at /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

For each configuration it builds the reference model (with the `timm` shim in
tests/golden/timm_shim), fills it with name-keyed deterministic weights
(`esmstereo_b200.weights.fill_deterministic`), calibrates BN running statistics with one train-mode
forward at momentum 1.0, runs the eval forward, and stores in
tests/golden/<name>.npz: the calibrated BN buffers (so the full state_dict can be rebuilt anywhere
from names alone) and the reference's stage outputs (cost volume checksums, stem/agg samples, cost,
top-2 indices, initial and final disparities, confidence).  tests/test_oracle_golden.py replays them
against oracle/esm_oracle.py; tests/test_gpu_parity.py against the CUDA path.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
sys.path.insert(1, "/root/reference")

from esmstereo_b200.weights import fill_deterministic, synthetic_pair  # noqa: E402

CONFIGS = {
    # name: (model, gwc, norm_corr, backbone, cv_scale, H, W)
    "cv4_gwc": ("ESMStereo", True, False, "efficientnet_b2", 4, 64, 128),
    "cv4_gwc_wide": ("ESMStereo", True, False, "efficientnet_b2", 4, 96, 224),
    "cv4_ncorr": ("ESMStereo", False, True, "efficientnet_b2", 4, 64, 128),
    "cv8_gwc": ("ESMStereo", True, False, "efficientnet_b2", 8, 64, 128),
    "cv16_gwc": ("ESMStereo", True, False, "mobilenetv2_100", 16, 96, 160),
    "cv16_ncorr": ("ESMStereo", False, True, "mobilenetv2_100", 16, 96, 160),
    "conf16_gwc": ("ESMStereo_confidence", True, False, "mobilenetv2_100", 16, 96, 160),
}


# BASELINE.json shapes, run through the real reference too; only samples of the big tensors are stored
CONFIGS_FULL = {
    "kitti_cv4_gwc": ("ESMStereo", True, False, "efficientnet_b2", 4, 384, 1248),
    "conf16_gwc_full": ("ESMStereo_confidence", True, False, "mobilenetv2_100", 16, 992, 1472),
}


def sample(t: torch.Tensor, limit: int = 20000) -> np.ndarray:
    """Deterministic strided subsample of a big tensor (keeps fixtures small)."""
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // limit)
    return flat[::step].to(torch.float32).numpy().copy()


def run(name, cfg, sampled=False):
    model_name, gwc, ncorr, backbone, s, H, W = cfg
    models = importlib.import_module("models")
    mod = importlib.import_module("models." + model_name)  # the module, not the re-bound class
    cls = models.__models__[model_name]
    torch.manual_seed(0)
    if model_name == "ESMStereo_confidence":
        net = cls(192, gwc, ncorr, backbone, s, torch.device("cpu"))
    else:
        net = cls(192, gwc, ncorr, backbone, s)
    sd = fill_deterministic(net.state_dict(), seed=0)
    net.load_state_dict(sd)

    left, right = synthetic_pair(1, H, W, shift=23 if sampled else 7, seed=0)
    # --- BN calibration: train mode, momentum 1.0
    for m in net.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = 1.0
    net.train()
    with torch.no_grad():
        if model_name == "ESMStereo":
            net(left, right, True)
        else:
            net(left, right)
    net.eval()

    cap = {}
    orig_gwc, orig_nc, orig_topk = mod.build_gwc_volume, mod.build_norm_correlation_volume, mod.regression_topk

    def cap_gwc(a, b, d, g):
        cap["match_left"], cap["match_right"] = a, b
        v = orig_gwc(a, b, d, g)
        cap["volume"] = v
        return v

    def cap_nc(a, b, d):
        cap["match_left"], cap["match_right"] = a, b
        v = orig_nc(a, b, d)
        cap["volume"] = v
        return v

    def cap_topk(cost, samples, k):
        _, ind = cost.sort(1, True)
        cap["top2_idx"] = ind[:, :k]
        return orig_topk(cost, samples, k)

    mod.build_gwc_volume, mod.build_norm_correlation_volume, mod.regression_topk = cap_gwc, cap_nc, cap_topk
    hooks = []

    def hook(key):
        def fn(_m, _i, o):
            cap[key] = o
        return fn

    for key, attr in (("stem", "group_stem" if gwc else "corr_stem"), ("agg", "agg"),
                      ("cost", "aggregation_out")):
        hooks.append(getattr(net, attr).register_forward_hook(hook(key)))
    up_in = {}

    def up_hook(_m, i, o):
        up_in["init_pred"], up_in["scales"] = i[-1], o

    hooks.append(net.upsample_module.register_forward_hook(up_hook))
    if model_name == "ESMStereo_confidence":
        cn = net.confidence_net

        def conf_hook(_m, i, o):
            cap["conf_init"], cap["conf_4"] = i[1], o

        hooks.append(cn.conf_up4.register_forward_hook(conf_hook))
        hooks.append(cn.embed_conv1.register_forward_hook(hook("conf_embed1_raw")))
        hooks.append(cn.embed_bn2.register_forward_hook(hook("conf_embed2_bn")))
    try:
        with torch.no_grad():
            if model_name == "ESMStereo":
                out = net(left, right, False)[-1]
                net_train_out = net(left, right, True)
                conf = None
            elif model_name == "ESMStereo_trt":
                out, conf, net_train_out = net(left, right), None, None
            else:
                out, conf = net(left, right)
                net_train_out = None
    finally:
        mod.build_gwc_volume, mod.build_norm_correlation_volume, mod.regression_topk = orig_gwc, orig_nc, orig_topk
        for h in hooks:
            h.remove()

    sd = net.state_dict()
    blob = {}
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            blob["bn/" + k] = v.numpy().astype(np.float32)
    vol = cap["volume"]
    blob["match_left_sample"] = sample(cap["match_left"])
    blob["match_right_sample"] = sample(cap["match_right"])
    blob["volume_sample"] = sample(vol)
    blob["volume_sum_over_hw"] = vol.sum((3, 4)).numpy().astype(np.float64)  # [B,G,D] checksum
    blob["volume_abs_sum"] = np.array(vol.abs().double().sum().item())
    blob["stem_sample"] = sample(cap["stem"])
    blob["agg_sample"] = sample(cap["agg"])
    if sampled:
        # full-size configuration: strided samples instead of whole tensors (fixtures stay < 1 MB)
        blob["cost_sample"] = sample(cap["cost"])
        blob["cost_absmax"] = np.array(cap["cost"].abs().max().item())
        blob["init_pred"] = up_in["init_pred"].numpy().astype(np.float32)
        if "top2_idx" in cap:
            blob["top2_idx"] = cap["top2_idx"].numpy().astype(np.int16)
        blob["disp_q"] = out[:, ::4, ::4].numpy().astype(np.float32)
        blob["disp_mean"] = np.array(out.double().mean().item())
        if conf is not None:
            blob["conf_q"] = conf[:, ::4, ::4].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        keys = {k: list(v.shape) for k, v in sd.items()}
        with open(os.path.join(HERE, name + ".keys.json"), "w") as f:
            json.dump({"config": list(cfg), "state_dict": keys}, f, indent=0, sort_keys=True)
        print(name, "ok (sampled): disp mean %.4f" % float(out.mean()))
        return
    blob["cost"] = cap["cost"].squeeze(1).numpy().astype(np.float32)
    blob["init_pred"] = up_in["init_pred"].numpy().astype(np.float32)
    if "top2_idx" in cap:
        blob["top2_idx"] = cap["top2_idx"].numpy().astype(np.int16)
    for i, t in enumerate(up_in["scales"]):
        blob["up_scale_%d" % i] = t.numpy().astype(np.float32)  # before squeeze and *4
    blob["disp"] = out.numpy().astype(np.float32)
    if net_train_out is not None:
        for i, t in enumerate(net_train_out):
            blob["train_out_%d" % i] = t.numpy().astype(np.float32)
    if conf is not None:
        blob["conf"] = conf.numpy().astype(np.float32)
        blob["conf_init"] = cap["conf_init"].numpy().astype(np.float32)
        blob["conf_4"] = cap["conf_4"].numpy().astype(np.float32)
        blob["conf_embed2_bn"] = cap["conf_embed2_bn"].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
    keys = {k: list(v.shape) for k, v in sd.items()}
    with open(os.path.join(HERE, name + ".keys.json"), "w") as f:
        json.dump({"config": list(cfg), "state_dict": keys}, f, indent=0, sort_keys=True)
    cost = cap["cost"]
    print("%-14s disp mean %.4f  cost std %.4f  size %.0f KB" % (
        name, out.mean().item(), cost.std().item(),
        os.path.getsize(os.path.join(HERE, name + ".npz")) / 1024))


if __name__ == "__main__":
    only = sys.argv[1:]
    for n, c in CONFIGS.items():
        if only and n not in only:
            continue
        run(n, c)
    for n, c in CONFIGS_FULL.items():
        if only and n not in only:
            continue
        if not only:
            continue  # the full-size ones take a minute each: only on request (python make_golden.py kitti_cv4_gwc conf16_gwc_full)
        run(n, c, sampled=True)
