"""Per-layer timing of the flat tcgen05 engine (conv_pf) against esm_conv_f32's autotuned choice, KITTI shapes."""
import os
import sys
import torch
sys.path.insert(0, ".")
from esmstereo_b200 import ops

dev = "cuda"
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=2, do_flush=False):
    """Device time per call: the call is captured into a CUDA graph holding 5 copies and replayed."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(5):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * iters)


def layer(cin, cout, k, nd, shape, stride=1, transposed=False, out="pf"):
    ks = (k,) * nd
    w = torch.randn(*(((cin, cout) if transposed else (cout, cin)) + ks), device=dev) * 0.05
    bn = (torch.ones(cout, device=dev), torch.zeros(cout, device=dev), torch.zeros(cout, device=dev), torch.ones(cout, device=dev), 1e-5)
    x = torch.randn(1, cin, *shape, device=dev)
    pad = 1 if (transposed or k == 3) else 0
    pc_old = ops.pack_conv(w, stride, pad, transposed, None, bn)
    pc_new = ops.pack_conv_pf(w, [cin], stride, transposed, None, bn)
    pf = ops.to_pf(x)
    t_old = timeit(lambda: ops.conv(x, pc_old, "gelu"))
    t_new = timeit(lambda: ops.conv_pf(pf, pc_new, "gelu", out=out))
    t_cvt = timeit(lambda: ops.to_pf(x))
    fl = 2.0 * cin * cout * k ** nd
    vox = 1
    for v in shape:
        vox *= v
    fl *= vox if (transposed or stride == 1) else vox / stride ** nd
    print("%-6s%dd %3d->%3d k%d s%d %-16s old %6.1f us | tcf %6.1f us (%5.1f TFLOP/s) | to_pf %5.1f us" % (
        "deconv" if transposed else "conv", nd, cin, cout, k, stride, "x".join(map(str, shape)), t_old, t_new, fl / t_new / 1e6, t_cvt), flush=True)


which = sys.argv[1:]
L = [
    (32, 32, 3, 2, (192, 624)), (64, 32, 3, 2, (192, 624)), (32, 16, 3, 2, (192, 624)), (96, 32, 1, 2, (192, 624)),
    (32, 32, 3, 2, (96, 312)), (80, 32, 3, 2, (96, 312)), (96, 64, 3, 2, (96, 312)), (48, 48, 3, 2, (96, 312)), (112, 32, 1, 2, (96, 312)),
    (32, 32, 3, 2, (48, 156)), (96, 96, 3, 2, (48, 156)), (160, 32, 1, 2, (48, 156)), (32, 32, 3, 2, (24, 78)), (240, 240, 3, 2, (24, 78)),
    (24, 24, 3, 3, (24, 48, 156)), (48, 24, 1, 3, (24, 48, 156)), (40, 40, 3, 3, (12, 24, 78)), (80, 40, 1, 3, (12, 24, 78)), (72, 72, 3, 3, (6, 12, 39)),
    (8, 8, 3, 3, (48, 96, 312)), (24, 8, 3, 3, (24, 48, 156)),
]
if which:
    for spec in which:
        v = spec.split(",")
        layer(int(v[0]), int(v[1]), int(v[2]), int(v[3]), tuple(int(t) for t in v[4].split("x")), int(v[5]) if len(v) > 5 else 1, len(v) > 6 and v[6] == "T")
    sys.exit(0)
for a in L:
    layer(*a)
S = [
    (32, 32, 3, 2, (192, 624), 2), (32, 32, 3, 2, (96, 312), 2), (32, 32, 3, 2, (48, 156), 2), (32, 48, 3, 2, (192, 624), 2),
    (8, 24, 3, 3, (48, 96, 312), 2), (24, 40, 3, 3, (24, 48, 156), 2), (40, 72, 3, 3, (12, 24, 78), 2),
]
for a in S:
    layer(*a)
T = [
    (32, 32, 4, 2, (96, 312), 2, True), (32, 32, 4, 2, (48, 156), 2, True), (32, 32, 4, 2, (24, 78), 2, True),
    (72, 40, 4, 3, (6, 12, 39), 2, True), (40, 24, 4, 3, (12, 24, 78), 2, True), (208, 120, 4, 2, (12, 39), 2, True), (240, 48, 4, 2, (24, 78), 2, True),
]
for a in T:
    layer(*a)
