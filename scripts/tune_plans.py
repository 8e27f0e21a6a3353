"""Tune the conv engine plans once on a B200 and write them to esmstereo_b200/plans/<name>.txt.

    ESM_PLANS=0 ESM_AUTOTUNE=1 python scripts/tune_plans.py [out.txt] [--only 0,3]     # --only: indices into RUNS

Every configuration the repo ships a measurement or a test for is run once in eager mode (each new layer shape is then
timed on the device by esm_conv_f32: FP32-pipe tilings, resident / streamed tcgen05, pointwise), and the chosen engine
and tiling per shape are exported as text.  Loaded at import (esmstereo_b200/_lib.py), the file pins those choices:
no timing, no synchronisation and the same rounding in every later process."""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ESM_PLANS", "0")
os.environ.setdefault("ESM_BACKBONE", "standin")
os.environ.setdefault("ESM_AUTOTUNE", "1")
from esmstereo_b200 import __models__, _lib  # noqa: E402
from esmstereo_b200.weights import fill_deterministic, synthetic_pair  # noqa: E402

RUNS = [
    # model, gwc, backbone, cv, batch, H, W
    ("ESMStereo", True, "efficientnet_b2", 4, 1, 384, 1248),
    ("ESMStereo", True, "efficientnet_b2", 4, 1, 256, 512),
    ("ESMStereo", True, "efficientnet_b2", 4, 1, 128, 256),
    ("ESMStereo", True, "efficientnet_b2", 4, 8, 544, 960),
    ("ESMStereo", True, "efficientnet_b2", 4, 2, 544, 960),
    ("ESMStereo", True, "efficientnet_b2", 4, 1, 544, 960),
    ("ESMStereo", True, "efficientnet_b2", 4, 2, 384, 1248),
    ("ESMStereo", True, "efficientnet_b2", 4, 4, 384, 1248),
    ("ESMStereo", True, "efficientnet_b2", 4, 8, 384, 1248),
    ("ESMStereo", True, "efficientnet_b2", 4, 16, 384, 1248),
    ("ESMStereo", True, "efficientnet_b2", 8, 1, 384, 1248),
    ("ESMStereo", True, "mobilenetv2_100", 16, 1, 384, 1248),
    ("ESMStereo_confidence", True, "mobilenetv2_100", 16, 1, 992, 1472),
    ("ESMStereo", False, "efficientnet_b2", 4, 1, 384, 1248),
]


def main():
    argv = list(sys.argv[1:])
    runs = RUNS
    if "--only" in argv:
        i = argv.index("--only")
        runs = [RUNS[int(t)] for t in argv[i + 1].split(",")]
        del argv[i:i + 2]
    out = argv[0] if argv else os.path.join(_lib.PLANS_DIR, "b200.txt")
    for name, gwc, backbone, cv, B, H, W in runs:
        with contextlib.redirect_stdout(io.StringIO()):
            m = __models__[name](192, gwc, not gwc, backbone, cv)
        m.load_state_dict(fill_deterministic(m.state_dict()))
        m = m.cuda().eval()
        l, r = [t.cuda() for t in synthetic_pair(B, H, W, shift=23, seed=1)]
        n0 = _lib.lib().esm_conv_tuned_calls()
        with torch.no_grad():
            for train_status in ((False, True) if name == "ESMStereo" else (None,)):
                if name == "ESMStereo_confidence":
                    m(l, r)
                else:
                    m(l, r, train_status=train_status)
        torch.cuda.synchronize()
        print("%-22s cv%-2d %s batch %2d %4dx%-4d: %3d shapes tuned" % (name, cv, "gwc" if gwc else "ncorr", B, H, W,
                                                                        _lib.lib().esm_conv_tuned_calls() - n0), flush=True)
        del m
        torch.cuda.empty_cache()
    # the golden / full-size test configurations (tests/test_gpu_model.py) use the same layer shapes at 64x128 ... 96x224
    text = _lib.export_plans()
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        f.write("# esm_conv_f32 engine plans tuned on %s (%d SMs) by scripts/tune_plans.py; format: include/esm_b200.h\n"
                % (torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count))
        f.write(text)
    print("wrote %d plans to %s" % (text.count("\n"), out))


if __name__ == "__main__":
    main()
