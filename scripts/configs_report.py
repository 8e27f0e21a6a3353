"""Throughput of the other BASELINE.json configurations (the bench line is configs[1]): SceneFlow-shaped 544x960 at
batch 1 / 8 (configs[2]), KITTI-shaped batch sweep (configs[3], one GPU's share), ESMStereo_confidence 992x1472 cv16
(configs[4]), and the cv8 / cv16 variants at KITTI shape.  CUDA-graph replay, CUDA events, 10 steps after 3 warm-ups."""
import contextlib, io, os, sys
os.environ.setdefault("ESM_BACKBONE", "standin")
import torch
sys.path.insert(0, ".")
from esmstereo_b200 import __models__, GraphedStereo
from esmstereo_b200.weights import fill_deterministic, synthetic_pair


def build(name, backbone, s):
    with contextlib.redirect_stdout(io.StringIO()):
        m = __models__[name](192, True, False, backbone, s)
    m.load_state_dict(fill_deterministic(m.state_dict()))
    return m.cuda().eval()


def timed(m, B, H, W, conf=False):
    l, r = [t.cuda() for t in synthetic_pair(B, H, W, shift=23, seed=3)]
    g = GraphedStereo(m, tuple(l.shape)) if conf else GraphedStereo(m, tuple(l.shape), train_status=False)
    for _ in range(3):
        g(l, r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = g(l, r)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    return ms, B / ms * 1e3


if __name__ == "__main__":
    m4 = build("ESMStereo", "efficientnet_b2", 4)
    for B in (1, 2, 4, 8):
        print("cv4  384x1248 batch %d: %.2f ms/step, %.1f pairs/s" % ((B,) + timed(m4, B, 384, 1248)), flush=True)
    for B in (1, 8):
        print("cv4  544x960  batch %d: %.2f ms/step, %.1f pairs/s" % ((B,) + timed(m4, B, 544, 960)), flush=True)
    del m4
    for s in (8, 16):
        ms = build("ESMStereo", "efficientnet_b2" if s == 8 else "mobilenetv2_100", s)
        print("cv%-2d 384x1248 batch 1: %.2f ms/step, %.1f pairs/s" % ((s,) + timed(ms, 1, 384, 1248)), flush=True)
        del ms
    try:
        mc = build("ESMStereo_confidence", "mobilenetv2_100", 16)
        print("conf 992x1472 batch 1: %.2f ms/step, %.1f pairs/s" % timed(mc, 1, 992, 1472, conf=True), flush=True)
    except Exception as e:  # GraphedStereo signature differences are reported, not hidden
        print("confidence config failed:", repr(e))
    print("peak mem %.2f GB" % (torch.cuda.max_memory_allocated() / 1e9))
