"""Top-2 flips / EPE of the SceneFlow-shape test case (544x960, batch 2, seed 1) per engine policy."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    from tests.helpers import rel_err
    from tests.test_gpu_model import _full_size, _flip_fraction
    m, orc, want, outs, _, _ = _full_size("ESMStereo", True, "efficientnet_b2", 4, 2, 544, 960, seed=1)
    cap = m.capture
    print("PARITY " + json.dumps({"cost": rel_err(cap["cost"].cpu().numpy(), want["cost"].numpy()), "flips": _flip_fraction(cap, want),
                                  "epe": float((outs[0].cpu() - want["disp"]).abs().mean())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for label, env in (("fp32", {"ESM_TC": "0", "ESM_PW_OFF": "1"}), ("resident-only", {"ESM_TCG_OFF": "1"}),
                           ("streamed on 2D only", {"ESM_TCG_DIMS": "2"}), ("streamed on 3D only", {"ESM_TCG_DIMS": "3"}), ("default", {})):
            out = subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, **env), capture_output=True, text=True)
            lines = [l for l in out.stdout.splitlines() if l.startswith("PARITY ")]
            print(label, lines[-1] if lines else out.stderr[-300:], flush=True)
