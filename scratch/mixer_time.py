"""Per-launch time of one SMLayer at 96 x 312 x 16: fused kernel against the two half kernels (graph of 20 launches)."""
import sys

import torch

sys.path.insert(0, ".")
from esmstereo_b200 import layers  # noqa: E402

torch.manual_seed(0)
lay = layers.SMLayer(16, 7, 2).cuda().eval()
x = torch.randn(1, 16, 96, 312, device="cuda")
for fused in (True, False):
    lay.fused = fused
    for _ in range(3):
        lay(x, extra_residual=x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            y = lay(x, extra_residual=x)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("SMLayer fused=%s: %.2f us" % (fused, e0.elapsed_time(e1) * 1e3 / 20))
