"""One eager KITTI-shaped forward with the role-timer build (ESM_TC_PROFILE=1): every tcgen05 launch prints block 0's
per-warp completion time and barrier waits (diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ.setdefault("ESM_BACKBONE", "standin")
import bench  # noqa: E402

cfg = dict(bench.CONFIGS["B"])
model, sd = bench.build_weights(cfg)
model = model.cuda().eval()
l = torch.randn(1, 3, cfg["H"], cfg["W"], device="cuda")
r = torch.randn_like(l)
with torch.no_grad():
    for i in range(2):
        if i == 1:
            print("=== profiled forward ===", flush=True)
        model(l, r, train_status=False)
        torch.cuda.synchronize()
