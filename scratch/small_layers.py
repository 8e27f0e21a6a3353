"""Per-launch time (graph of 50 back-to-back launches, warm) of small 2D layers per engine (diagnostic).

    python scratch/small_layers.py            # timings
    ESM_TC_PROFILE=1 python scratch/small_layers.py prof   # role timers of the resident tcgen05 engine, one launch each
"""
import os
import sys

import torch

sys.path.insert(0, ".")
from esmstereo_b200 import ops  # noqa: E402
from scripts.prof_conv import layer  # noqa: E402


def timeit(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


CASES = [(32, 32, 3, (24, 78)), (32, 32, 3, (48, 156)), (32, 32, 3, (96, 312)), (32, 32, 3, (192, 624)), (16, 32, 3, (96, 312)),
         (80, 32, 3, (96, 312)), (32, 16, 1, (96, 312)), (128, 32, 1, (48, 156)), (16, 16, 3, (192, 624)), (48, 48, 3, (96, 312)),
         (32, 32, 1, (96, 312)), (160, 32, 1, (48, 156)), (112, 32, 1, (96, 312)), (96, 32, 1, (192, 624))]

if __name__ == "__main__":
    prof = "prof" in sys.argv
    for cin, cout, k, hw in CASES:
        x = torch.randn(1, cin, *hw, device="cuda")
        pc = layer(cin, cout, k, 2, pad=k // 2)
        if prof:
            os.environ["ESM_TC"], os.environ["ESM_TC_FORCE"] = "3", "1"
            print("== %d->%d k%d %s" % (cin, cout, k, hw), flush=True)
            ops.conv(x, pc, "gelu")
            torch.cuda.synchronize()
            continue
        row = []
        for name, env in (("fp32", {"ESM_TC": "0"}), ("tc", {"ESM_TC": "3", "ESM_TC_FORCE": "1"}), ("tcg", {"ESM_TC": "3", "ESM_TC_FORCE": "2"}),
                          ("pw", {"ESM_TC": "3", "ESM_TC_FORCE": "3"}), ("pinned", {"ESM_TC": "3"})):
            os.environ.pop("ESM_TC_FORCE", None)
            os.environ.update(env)
            try:
                row.append("%s %6.2f" % (name, timeit(lambda: ops.conv(x, pc, "gelu"))))
            except Exception as e:  # noqa: BLE001
                row.append("%s   err" % name)
        print("%3d->%-3d k%d %-10s : %s us/launch" % (cin, cout, k, hw, "   ".join(row)), flush=True)
