"""Per-stage GPU-vs-oracle error report (diagnostics; prints, never asserts)."""
import contextlib
import io
import sys

import torch

sys.path.insert(0, ".")
from esmstereo_b200 import GraphedStereo, __models__  # noqa: E402
from esmstereo_b200.weights import fill_deterministic, synthetic_pair  # noqa: E402
from oracle.esm_oracle import EsmOracle  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


def report(model_name, gwc, backbone, s, B, H, W, seed):
    print("=== %s gwc=%s cv%d %dx%dx%d seed %d" % (model_name, gwc, s, B, H, W, seed), flush=True)
    with contextlib.redirect_stdout(io.StringIO()):
        m = __models__[model_name](192, gwc, not gwc, backbone, s)
    sd = fill_deterministic(m.state_dict(), seed=seed)
    left, right = synthetic_pair(B, H, W, shift=23, seed=seed)
    conf = model_name == "ESMStereo_confidence"
    orc = EsmOracle(sd, 192, gwc, not gwc, backbone, s, confidence=conf)
    sd = orc.calibrate(left[:1], right[:1])
    want = orc(left, right)
    want64 = EsmOracle(sd, 192, gwc, not gwc, backbone, s, confidence=conf, dtype=torch.float64)(left, right)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.capture = {}
    if conf:
        m.confidence_net.capture = {}
    l, r = left.cuda(), right.cuda()
    out = m(l, r, False)[-1] if model_name == "ESMStereo" else m(l, r)
    cap = dict(m.capture)
    if conf:
        cap.update(m.confidence_net.capture)
        out, cf = out
        cap["conf"] = cf
    cap["disp"] = out
    for k in ["match_left", "stem", "agg", "cost", "init_pred", "disp", "conf_top7", "conf_feat", "conf_scale", "conf_embed",
              "conf_init", "conf_4", "conf"]:
        if k in cap and k in want:
            print("  %-11s gpu-vs-cpu32 rel %.3e | cpu32-vs-cpu64 rel %.3e | gpu-vs-cpu64 rel %.3e" % (
                k, rel(cap[k], want[k]), rel(want[k], want64[k]), rel(cap[k], want64[k])))
    if "top2_idx" in cap:
        g = cap["top2_idx"].cpu().long().sort(1).values
        for nm, w in (("cpu32", want), ("cpu64", want64)):
            mm = (g != w["top2_idx"].sort(1).values).any(1).float()
            print("  top2 mismatching pixels vs %s: %d of %d" % (nm, int(mm.sum()), mm.numel()))
        mm = (want["top2_idx"].sort(1).values != want64["top2_idx"].sort(1).values).any(1).float()
        print("  top2 mismatching pixels cpu32 vs cpu64: %d" % int(mm.sum()))
    for b in range(B):
        print("  image %d: EPE gpu-vs-cpu32 %.3e  cpu32-vs-cpu64 %.3e  gpu-vs-cpu64 %.3e" % (
            b, float((out[b].cpu() - want["disp"][b]).abs().mean()),
            float((want["disp"][b].double() - want64["disp"][b]).abs().mean()),
            float((out[b].cpu().double() - want64["disp"][b]).abs().mean())))
    # determinism: eager twice, graph replay
    m.capture = None
    if conf:
        m.confidence_net.capture = None
    run = (lambda: m(l, r, False)[-1]) if model_name == "ESMStereo" else (lambda: m(l, r)[0])
    a, b2 = run(), run()
    print("  eager run-to-run max diff %.3e" % float((a - b2).abs().max()))
    kw = {"train_status": False} if model_name == "ESMStereo" else {}
    g = GraphedStereo(m, tuple(l.shape), **kw)
    o = g(l, r)
    o = o[-1] if model_name == "ESMStereo" else o[0]
    torch.cuda.synchronize()
    print("  graph-vs-eager max diff %.3e" % float((o - a).abs().max()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["kitti", "sf", "conf"]
    if "kitti" in which:
        report("ESMStereo", True, "efficientnet_b2", 4, 1, 384, 1248, 0)
    if "sf" in which:
        report("ESMStereo", True, "efficientnet_b2", 4, 2, 544, 960, 1)
    if "conf" in which:
        report("ESMStereo_confidence", True, "mobilenetv2_100", 16, 1, 96, 160, 0)
        report("ESMStereo_confidence", True, "mobilenetv2_100", 16, 1, 992, 1472, 2)
