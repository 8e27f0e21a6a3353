"""Device time of the 3 -> 32 k3 s2 image-side layers at KITTI shape: dedicated kernel vs esm_conv_f32's engines."""
import os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scratch")
from esmstereo_b200 import ops
from _timing import timeit

x = torch.randn(2, 3, 384, 1248, device="cuda")
w = torch.randn(32, 3, 3, 3, device="cuda") * 0.2
bn = (torch.ones(32, device="cuda"), torch.zeros(32, device="cuda"), torch.zeros(32, device="cuda"), torch.ones(32, device="cuda"), 1e-5)
pc = ops.pack_conv(w, 2, 1, False, None, bn)
for act in ("relu6", "gelu"):
    print(act, "%.1f us" % timeit(lambda: ops.conv(x, pc, act)), flush=True)
