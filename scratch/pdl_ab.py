"""A/B of programmatic dependent launch on the whole KITTI-shaped forward in ONE process: one CUDA graph per mask,
replays interleaved (diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ.setdefault("ESM_BACKBONE", "standin")
import bench  # noqa: E402
from esmstereo_b200._lib import lib  # noqa: E402

cfg = dict(bench.CONFIGS["B"])
model, sd = bench.build_weights(cfg)
model = model.cuda().eval()
l = torch.randn(1, 3, cfg["H"], cfg["W"], device="cuda")
r = torch.randn_like(l)
masks = [int(a) for a in sys.argv[1:]] or [0, 62, 63, 2, 4, 8, 16, 32, 1]
graphs = {}
with torch.no_grad():
    for _ in range(3):
        ref = model(l, r, train_status=False)
    torch.cuda.synchronize()
    for m in masks:
        lib().esm_set_pdl(m)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = model(l, r, train_status=False)
        graphs[m] = (g, out)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {m: [] for m in masks}
for rep in range(5):
    for m in masks:
        g, out = graphs[m]
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[m].append(e0.elapsed_time(e1) / 20)
for m in masks:
    g, out = graphs[m]
    same = all(torch.equal(a, b) for a, b in zip(out, graphs[masks[0]][1])) if isinstance(out, (list, tuple)) else torch.equal(out, graphs[masks[0]][1])
    print("mask %2d: %.4f ms/step (min %.4f)  identical to mask %d: %s" % (m, sorted(res[m])[len(res[m]) // 2], min(res[m]), masks[0], same))
