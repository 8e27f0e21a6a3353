"""Per-layer accuracy of the conv engines against an fp64 reference (mean and max error relative to the output RMS).
One process per engine (plans are cached per process)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = {"c40_3d": (40, 40, 3, 3, 1, 1, False, (1, 40, 12, 24, 78)), "c24_3d": (24, 24, 3, 3, 1, 1, False, (1, 24, 24, 48, 156)),
         "s2_24_40": (24, 40, 3, 3, 2, 1, False, (1, 24, 24, 48, 156)), "d40_24": (40, 24, 4, 3, 2, 1, True, (1, 40, 12, 24, 78)),
         "c240_2d": (240, 240, 3, 2, 1, 1, False, (1, 240, 24, 78)), "c96_64_2d": (96, 64, 3, 2, 1, 1, False, (1, 96, 96, 312))}


def one():
    import torch
    import torch.nn.functional as F
    from esmstereo_b200 import ops
    res = {}
    for name, (cin, cout, k, nd, stride, pad, tr, shape) in CASES.items():
        g = torch.Generator().manual_seed(7)
        w = torch.randn(*(((cin, cout) if tr else (cout, cin)) + (k,) * nd), generator=g) * (2.0 / (cin * k ** nd)) ** 0.5
        x = torch.randn(*shape, generator=g).abs()  # post-GELU-like, mostly positive: a coherent (not zero-mean) sum
        fn = {(2, False): F.conv2d, (2, True): F.conv_transpose2d, (3, False): F.conv3d, (3, True): F.conv_transpose3d}[(nd, tr)]
        ref = fn(x.double(), w.double(), stride=stride, padding=pad)
        pc = ops.pack_conv(w.cuda(), stride, pad, tr, None, None)
        got = ops.conv(x.cuda(), pc, None).cpu().double()
        rms = ref.pow(2).mean().sqrt()
        res[name] = [float((got - ref).abs().mean() / rms), float((got - ref).abs().max() / rms), float(((got - ref) * ref.sign()).mean() / rms)]
    print("ACC " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for label, env in (("fp32", {"ESM_TC": "0", "ESM_PW_OFF": "1"}), ("resident", {"ESM_TC_FORCE": "1"}), ("streamed", {"ESM_TC_FORCE": "2"})):
            out = subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, **env), capture_output=True, text=True)
            lines = [l for l in out.stdout.splitlines() if l.startswith("ACC ")]
            if not lines:
                print(label, out.stderr[-300:]); continue
            d = json.loads(lines[-1][4:])
            print("%-9s" % label + "  ".join("%s mean %.1e max %.1e bias %+.1e" % (k, v[0], v[1], v[2]) for k, v in d.items()), flush=True)
