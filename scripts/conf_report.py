"""End-to-end confidence error at 992x1472 (BASELINE configs[4]) against the fp64 oracle, per engine policy."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    from oracle.esm_oracle import EsmOracle
    from tests.test_gpu_model import _full_size
    from esmstereo_b200 import _lib
    m, orc, want, outs, conf, (left, right) = _full_size("ESMStereo_confidence", True, "mobilenetv2_100", 16, 1, 992, 1472, seed=2)
    want64 = EsmOracle(orc.sd, 192, True, False, "mobilenetv2_100", 16, confidence=True, dtype=torch.float64)(left, right)
    d = (conf.cpu().double() - want64["conf"]).abs()
    c = (want["conf"].double() - want64["conf"]).abs()
    print("CONF " + json.dumps({"gpu_max": float(d.max()), "gpu_mean": float(d.mean()), "cpu32_max": float(c.max()), "cpu32_mean": float(c.mean()),
                                "epe": float((outs[0].cpu() - want["disp"]).abs().mean()),
                                "tc": int(_lib.lib().esm_tc_conv_launches()), "tcg": int(_lib.lib().esm_tcg_conv_launches())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for label, env in (("fp32", {"ESM_TC": "0"}), ("resident-only", {"ESM_TCG_OFF": "1"}), ("default", {})):
            out = subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, **env), capture_output=True, text=True)
            lines = [l for l in out.stdout.splitlines() if l.startswith("CONF ")]
            print(label, lines[-1] if lines else out.stderr[-500:], flush=True)
