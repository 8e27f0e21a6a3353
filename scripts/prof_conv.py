"""Run a handful of representative hot-path kernels at KITTI shapes (for ncu captures)."""
import sys

import torch

sys.path.insert(0, ".")
from esmstereo_b200 import ops  # noqa: E402


def layer(cin, cout, k, nd, stride=1, pad=1, transposed=False):
    ks = (k,) * nd
    w = torch.randn(*(((cin, cout) if transposed else (cout, cin)) + ks), device="cuda") * 0.05
    bn = (torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda"),
          torch.ones(cout, device="cuda"), 1e-5)
    return ops.pack_conv(w, stride, pad, transposed, None, bn)


def main():
    which = sys.argv[1:] or ["stem", "agg", "c24", "s2", "c2d", "dec1", "vol"]
    h, w, D = 96, 312, 48
    L, R = torch.randn(1, 64, h, w, device="cuda"), torch.randn(1, 64, h, w, device="cuda")
    reps = 3
    for _ in range(reps):
        if "vol" in which:
            ops.build_gwc_volume(L, R, D, 32)
        if "stem" in which:
            ops.conv([L, R], layer(32, 8, 3, 3), "gelu", gwc_disp=D)
        if "agg" in which:
            ops.conv(torch.randn(1, 8, D, h, w, device="cuda"), layer(8, 8, 3, 3), "gelu")
        if "s2" in which:
            ops.conv(torch.randn(1, 8, D, h, w, device="cuda"), layer(8, 24, 3, 3, stride=2), "gelu")
        if "c24" in which:
            ops.conv(torch.randn(1, 24, 24, 48, 156, device="cuda"), layer(24, 24, 3, 3), "gelu")
        if "c2d" in which:
            ops.conv(torch.randn(1, 32, 192, 624, device="cuda"), layer(32, 32, 3, 2), "gelu")
        if "dec1" in which:
            ops.conv(torch.randn(1, 24, 24, 48, 156, device="cuda"), layer(24, 1, 4, 3, stride=2, transposed=True), None)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
