"""BASELINE.json configs[2]: SceneFlow-shaped 544x960 pairs, batch 8, on one B200 (throughput path)."""
import contextlib, io, sys, time
import torch
sys.path.insert(0, ".")
from esmstereo_b200 import __models__, GraphedStereo
from esmstereo_b200.weights import fill_deterministic, synthetic_pair

with contextlib.redirect_stdout(io.StringIO()):
    m = __models__["ESMStereo"](192, True, False, "efficientnet_b2", 4)
m.load_state_dict(fill_deterministic(m.state_dict()))
m = m.cuda().eval()
for B in (1, 8):
    l, r = [t.cuda() for t in synthetic_pair(B, 544, 960, shift=23, seed=3)]
    g = GraphedStereo(m, tuple(l.shape), train_status=False)
    for _ in range(3):
        g(l, r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        out = g(l, r)[-1]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("544x960 batch %d: %.2f ms/step, %.1f pairs/s, out %s finite=%s, peak mem %.2f GB" % (
        B, ms, B / ms * 1e3, tuple(out.shape), bool(torch.isfinite(out).all()), torch.cuda.max_memory_allocated() / 1e9))
