"""Parity of the full KITTI-shape forward (384x1248, cv4 gwc) against the CPU oracle for each tensor-core
policy: ESM_TC=0 (FP32 pipe only), 3 (default: split-TF32 tcgen05 where it wins the timing), 1 (single-pass
TF32 fast mode).  Prints one line per policy: cost-volume relative error, top-2 index flip fraction, EPE.

    python scripts/parity_report.py            # each policy runs in its own process (plans are cached per process)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    from tests.helpers import rel_err
    from tests.test_gpu_model import _full_size, _flip_fraction
    from esmstereo_b200 import _lib
    m, orc, want, outs, _, _ = _full_size("ESMStereo", True, "efficientnet_b2", 4, 1, 384, 1248)
    cap = m.capture
    res = {"tc": os.environ.get("ESM_TC", "3"),
           "cost_rel_err": rel_err(cap["cost"].cpu().numpy(), want["cost"].numpy()),
           "top2_flip_fraction": _flip_fraction(cap, want),
           "epe_px": float((outs[0].cpu() - want["disp"]).abs().mean()),
           "max_abs_px": float((outs[0].cpu() - want["disp"]).abs().max()),
           "tc_conv_launches": int(_lib.lib().esm_tc_conv_launches())}
    print("PARITY " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for tc in ("0", "3", "1"):
            env = dict(os.environ, ESM_TC=tc)
            out = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
            lines = [l for l in out.stdout.splitlines() if l.startswith("PARITY ")]
            print(lines[-1] if lines else "FAILED tc=%s: %s" % (tc, out.stderr[-400:]), flush=True)
