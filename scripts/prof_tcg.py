"""Time (CUDA events, L2 flushed) the streamed-weight tcgen05 engine (conv_tcg.cu) against the other engines on
the KITTI-shape layers it targets.

    python scripts/prof_tcg.py [c40 c72 s2_8_24 ...]       # timings: fp32 pipe, resident tcgen05, streamed tcgen05
    python scripts/prof_tcg.py --one c40                   # one forced streamed launch (role profiler / ncu)
"""
import os
import sys

import torch

sys.path.insert(0, ".")
from esmstereo_b200 import ops  # noqa: E402
from scripts.prof_conv import layer  # noqa: E402

CASES = {
    # name: (cin, cout, k, nd, stride, pad, transposed, input shape, GFLOP)
    "c40": (40, 40, 3, 3, 1, 1, False, (1, 40, 12, 24, 78)),
    "c72": (72, 72, 3, 3, 1, 1, False, (1, 72, 6, 12, 39)),
    "s2_8_24": (8, 24, 3, 3, 2, 1, False, (1, 8, 48, 96, 312)),
    "s2_24_40": (24, 40, 3, 3, 2, 1, False, (1, 24, 24, 48, 156)),
    "s2_40_72": (40, 72, 3, 3, 2, 1, False, (1, 40, 12, 24, 78)),
    "d72_40": (72, 40, 4, 3, 2, 1, True, (1, 72, 6, 12, 39)),
    "d40_24": (40, 24, 4, 3, 2, 1, True, (1, 40, 12, 24, 78)),
    "c240": (240, 240, 3, 2, 1, 1, False, (1, 240, 24, 78)),
    "d208_120": (208, 120, 4, 2, 2, 1, True, (1, 208, 12, 39)),
    "d240_48": (240, 48, 4, 2, 2, 1, True, (1, 240, 24, 78)),
    "d96_24": (96, 24, 4, 2, 2, 1, True, (1, 96, 48, 156)),
    "s2_32_48": (32, 48, 3, 2, 2, 1, False, (1, 32, 192, 624)),
}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    one = "--one" in sys.argv
    which = args or list(CASES)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in which:
        cin, cout, k, nd, stride, pad, tr, shape = CASES[name]
        pc = layer(cin, cout, k, nd, stride=stride, pad=pad, transposed=tr)
        x = torch.randn(*shape, device="cuda")
        run = lambda: ops.conv(x, pc, "gelu")
        for mode in (["tcg"] if one else ["fp32", "tc", "tcg"]):
            os.environ["ESM_TC"] = "0" if mode == "fp32" else "3"
            os.environ["ESM_TC_FORCE"] = "2" if mode == "tcg" else "1"
            run()
            torch.cuda.synchronize()
            if one:
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
            print("%-9s %-5s %8.1f us" % (name, mode, ts[len(ts) // 2]), flush=True)
    print("ok")


if __name__ == "__main__":
    main()
