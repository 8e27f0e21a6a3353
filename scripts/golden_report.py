"""Parity of every golden configuration (tests/golden) under each engine policy, one process per policy:
cost relative error, top-2 index mismatch fraction, EPE -- to see what a policy change does to the gates.

    python scripts/golden_report.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    from tests.helpers import GOLDEN_NAMES, golden_blob, golden_config, golden_inputs, golden_state_dict, rel_err, sample
    from tests.test_gpu_model import build, run
    from esmstereo_b200 import _lib
    for name in GOLDEN_NAMES:
        cfg, blob = golden_config(name), golden_blob(name)
        m = build(cfg["model"], cfg["gwc"], cfg["norm_correlation"], cfg["backbone"], cfg["cv_scale"], golden_state_dict(name))
        m.capture = {}
        left, right = golden_inputs(name)
        outs, conf = run(m, cfg["model"], left.cuda(), right.cuda())
        cap = m.capture
        res = {"name": name, "match": rel_err(sample(cap["match_left"]), blob["match_left_sample"]),
               "agg": rel_err(sample(cap["agg"]), blob["agg_sample"]), "cost": rel_err(cap["cost"].cpu().numpy(), blob["cost"])}
        if "top2_idx" in blob.files:
            res["top2_mismatch"] = float((np.sort(cap["top2_idx"].cpu().numpy(), 1) != np.sort(blob["top2_idx"].astype(np.int32), 1)).mean())
        res["epe"] = float(np.abs(outs[0].cpu().numpy() - blob["disp"]).mean())
        res["tc"], res["tcg"] = int(_lib.lib().esm_tc_conv_launches()), int(_lib.lib().esm_tcg_conv_launches())
        print("GOLDEN " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for label, env in (("fp32", {"ESM_TC": "0"}), ("resident-only", {"ESM_TCG_OFF": "1"}), ("default", {})):
            out = subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, **env), capture_output=True, text=True)
            print("== " + label, flush=True)
            for l in out.stdout.splitlines():
                if l.startswith("GOLDEN "):
                    d = json.loads(l[7:])
                    print("  %-12s match %.1e agg %.1e cost %.1e top2 %s epe %.4f  (tc %d, tcg %d launches)" % (
                        d["name"], d["match"], d["agg"], d["cost"], ("%.1e" % d["top2_mismatch"]) if "top2_mismatch" in d else "-", d["epe"], d["tc"], d["tcg"]), flush=True)
            if out.returncode:
                print(out.stderr[-600:])
