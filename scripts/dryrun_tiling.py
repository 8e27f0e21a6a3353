"""CPU dry run: record every hot-path convolution the oracle executes at a given configuration and ask
libesm_b200 to validate + tile it (fake pointers; without a GPU the launch itself returns ESM_ERR_CUDA,
which is the success signal here -- ESM_ERR_ARG means the shape has no tiling / is rejected)."""
import contextlib
import ctypes as C
import io
import sys

import torch

sys.path.insert(0, ".")
from esmstereo_b200 import __models__, _lib  # noqa: E402
from esmstereo_b200.weights import fill_deterministic, synthetic_pair  # noqa: E402
from oracle.esm_oracle import EsmOracle  # noqa: E402

HOT = ("group_stem", "corr_stem", "agg.", "aggregation_out", "upsample_module", "confidence_net")


def record(model_name, gwc, backbone, s, H, W, B=1):
    with contextlib.redirect_stdout(io.StringIO()):
        m = __models__[model_name](192, gwc, not gwc, backbone, s)
    sd = fill_deterministic(m.state_dict())
    orc = EsmOracle(sd, 192, gwc, not gwc, backbone, s, confidence=model_name == "ESMStereo_confidence")
    rec = []
    orig = orc.conv

    def conv(x, wkey, stride=1, pad=1, deconv=False, bias=None, groups=1):
        y = orig(x, wkey, stride, pad, deconv, bias, groups)
        if wkey.startswith(HOT) and groups == 1:
            rec.append((wkey, tuple(x.shape), tuple(orc.w(wkey).shape), stride, pad, deconv, tuple(y.shape)))
        return y

    orc.conv = conv
    orc(*synthetic_pair(B, H, W))
    return rec


def dry(rec):
    L = _lib.lib()
    bad = 0
    for wkey, xs, ws, stride, pad, deconv, ys in rec:
        nd = len(xs) - 2
        d = _lib.EsmConv()
        Cin, Cout = (ws[0], ws[1]) if deconv else (ws[1], ws[0])
        if "embed_conv2" in wkey or "conf_spx.weight" in wkey:
            continue  # handled by dedicated fused kernels
        d.src[0].ptr, d.src[0].C = 16, Cin
        d.nsrc = 1
        d.B, d.Cin = xs[0], Cin
        d.Din, d.Hin, d.Win = (xs[2], xs[3], xs[4]) if nd == 3 else (1, xs[2], xs[3])
        d.Cout = Cout
        d.Dout, d.Hout, d.Wout = (ys[2], ys[3], ys[4]) if nd == 3 else (1, ys[2], ys[3])
        k = ws[2:]
        d.kd, d.kh, d.kw = k if nd == 3 else (1,) + tuple(k)
        d.stride = stride
        d.pd, d.ph, d.pw = (pad, pad, pad) if nd == 3 else (0, pad, pad)
        d.transposed = int(deconv)
        d.weight, d.out, d.out_scale = 16, 16, 1.0
        rc = L.esm_conv_f32(C.byref(d), None)
        if rc == -1:
            bad += 1
            print("NO TILING", wkey, xs, ws, L.esm_last_error().decode())
    return bad


if __name__ == "__main__":
    total = 0
    for cfg in [("ESMStereo", True, "efficientnet_b2", 4, 384, 1248), ("ESMStereo", True, "efficientnet_b2", 4, 64, 128),
                ("ESMStereo", True, "efficientnet_b2", 4, 544, 960), ("ESMStereo", True, "efficientnet_b2", 8, 384, 1248),
                ("ESMStereo", True, "efficientnet_b2", 8, 64, 128), ("ESMStereo", False, "efficientnet_b2", 4, 96, 224),
                ("ESMStereo_confidence", True, "mobilenetv2_100", 16, 992, 1472),
                ("ESMStereo_confidence", True, "mobilenetv2_100", 16, 96, 160)]:
        rec = record(*cfg)
        b = dry(rec)
        total += b
        print(cfg, "convs:", len(rec), "rejected:", b)
    sys.exit(1 if total else 0)
