"""Dump the CUDA graph of one KITTI-shaped forward (DOT, verbose) to count node kinds and programmatic edges (diagnostic)."""
import os
import re
import sys
from collections import Counter

import torch

sys.path.insert(0, ".")
os.environ.setdefault("ESM_BACKBONE", "standin")
import bench  # noqa: E402

cfg = dict(bench.CONFIGS["B"])
model, sd = bench.build_weights(cfg)
model = model.cuda().eval()
l = torch.zeros(1, 3, cfg["H"], cfg["W"], device="cuda")
r = torch.zeros_like(l)
with torch.no_grad():
    for _ in range(3):
        model(l, r, train_status=False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(keep_graph=True)
    g.enable_debug_mode()
    with torch.cuda.graph(g):
        out = model(l, r, train_status=False)
    out_path = os.path.abspath("gpurun_out/graph.dot")
    g.debug_dump(out_path)
    print("exists", os.path.exists(out_path), os.listdir("gpurun_out"))
txt = open(out_path).read()
print("bytes", len(txt))
kinds = Counter(re.findall(r'label="\{?\s*([A-Z_a-z ]+?)\s*[\|\\]', txt))
print(kinds.most_common(12))
print("edges", txt.count("->"), "programmatic-ish", len(re.findall(r"(?i)programmatic", txt)))
# keep the dump small: strip it to node ids, kernel names and edges
keep = [ln for ln in txt.splitlines() if "->" in ln or "label" in ln]
open("gpurun_out/graph_small.dot", "w").write("\n".join(x[:400] for x in keep))
os.remove(out_path)
