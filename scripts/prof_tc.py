"""Time (CUDA events, L2 flushed) and optionally profile the tensor-core conv path against the FP32-pipe
kernels on the KITTI-shape layers.

    python scripts/prof_tc.py [stem agg c24 c2d ...]          # timings, both paths
    ncu --set full --import-source on --profile-from-start off -k regex:tc_conv ... python scripts/prof_tc.py --ncu stem
"""
import os
import sys

import torch

sys.path.insert(0, ".")
from esmstereo_b200 import ops  # noqa: E402
from scripts.prof_conv import layer  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ncu = "--ncu" in sys.argv
    which = args or ["stem", "agg", "c24", "c2d", "c2d64", "c96"]
    h, w, D = 96, 312, 48
    L, R = torch.randn(1, 64, h, w, device="cuda"), torch.randn(1, 64, h, w, device="cuda")
    cases = {
        "stem": (layer(32, 8, 3, 3), lambda pc: ops.conv([L, R], pc, "gelu", gwc_disp=D), 19.875),
        "agg": (layer(8, 8, 3, 3), torch.randn(1, 8, D, h, w, device="cuda"), 4.969),
        "c24": (layer(24, 24, 3, 3), torch.randn(1, 24, 24, 48, 156, device="cuda"), 5.590),
        "c2d": (layer(32, 32, 3, 2), torch.randn(1, 32, 192, 624, device="cuda"), 2.208),
        "c2d64": (layer(64, 32, 3, 2), torch.randn(1, 64, 192, 624, device="cuda"), 4.416),
        "c96": (layer(96, 64, 3, 2), torch.randn(1, 96, 96, 312, device="cuda"), 3.312),
    }
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in which:
        pc, x, gflop = cases[name]
        run = (lambda: x(pc)) if callable(x) else (lambda: ops.conv(x, pc, "gelu"))
        for mode in (["tc"] if ncu else ["fp32", "tc", "tc1"]):
            os.environ["ESM_TC"] = {"fp32": "0", "tc": "3", "tc1": "1"}[mode]
            os.environ["ESM_TC_FORCE"] = "1"
            run()
            torch.cuda.synchronize()
            if ncu:
                torch.cuda.profiler.start()
                run()
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
                continue
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            us = ts[len(ts) // 2]
            print("%-6s %-5s %8.1f us  %6.1f TFLOP/s" % (name, mode, us, gflop / us * 1e3 / 1e3), flush=True)
    print("ok")


if __name__ == "__main__":
    main()
