"""ncu target: the pointwise kernel on the hourglass's agg_1.0 (48 -> 24 k1 3D, two sources) and on ref4x's agg_1.0
(96 -> 32 k1 2D at 192 x 624, three sources) -- diagnostic."""
import os, sys
import torch
sys.path.insert(0, ".")
os.environ["ESM_TC_FORCE"] = "3"
from esmstereo_b200 import ops
from scripts.prof_conv import layer
a, b = torch.randn(1, 24, 24, 48, 156, device="cuda"), torch.randn(1, 24, 24, 48, 156, device="cuda")
pc3 = layer(48, 24, 1, 3, pad=0)
x1, x2, x3 = torch.randn(1, 32, 192, 624, device="cuda"), torch.randn(1, 32, 192, 624, device="cuda"), torch.randn(1, 32, 192, 624, device="cuda")
pc2 = layer(96, 32, 1, 2, pad=0)
def go():
    ops.conv([a, b], pc3, "gelu")
    ops.conv([x1, x2, x3], pc2, "gelu")
go(); torch.cuda.synchronize()
torch.cuda.profiler.start(); go(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok")
