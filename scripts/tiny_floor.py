"""Fixed per-launch cost of the conv kernel on a trivially small problem (diagnostic)."""
import sys, os
import torch
sys.path.insert(0, ".")
from esmstereo_b200 import ops
from scripts.prof_conv import layer

def timeit(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

for (cin, cout, k, hw, act) in [(8, 8, 1, (32, 32), None), (8, 8, 1, (32, 32), "gelu"), (8, 8, 3, (32, 32), "gelu"), (32, 32, 3, (24, 80), "gelu"),
                                (32, 16, 1, (96, 312), None), (32, 32, 3, (96, 312), "gelu")]:
    x = torch.randn(1, cin, *hw, device="cuda")
    pc = layer(cin, cout, k, 2, pad=k // 2)
    t = timeit(lambda: ops.conv(x, pc, act))
    print("conv %d->%d k%d %s act=%s: %.2f us/launch (ESM_NO_TMA=%s)" % (cin, cout, k, hw, act, t, os.environ.get("ESM_NO_TMA")))
y = torch.randn(1, 16, 96, 312, device="cuda")
print("torch add: %.2f us/launch" % timeit(lambda: y + 1.0))
