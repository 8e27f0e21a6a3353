"""First light of the flat tcgen05 engine: conv_pf vs a float64 torch convolution."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from esmstereo_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def check(name, got, want, tol=2e-5):
    err = (got.double().cpu() - want).abs().max().item() / max(want.abs().max().item(), 1e-9)
    print("%-50s rel err %.2e %s" % (name, err, "OK" if err < tol else "FAIL"), flush=True)
    return err < tol


def run(cin, cout, k, nd, shape, act="gelu", stride=1, transposed=False, nsrc=1, B=1):
    ks = (k,) * nd
    w = torch.randn(*(((cin, cout) if transposed else (cout, cin)) + ks)) * (0.5 / (cin * k ** nd) ** 0.5)
    bias = torch.randn(cout) * 0.1
    x = torch.randn(B, cin, *shape)
    conv = {2: (F.conv_transpose2d if transposed else F.conv2d), 3: (F.conv_transpose3d if transposed else F.conv3d)}[nd]
    want = conv(x.double(), w.double(), bias.double(), stride=stride, padding=(1 if (transposed or k == 3) else 0))
    if act == "gelu":
        want = F.gelu(want)
    split = [cin] if nsrc == 1 else [cin // 2 // 8 * 8, cin - cin // 2 // 8 * 8]
    xs = torch.split(x, split, 1)
    pfs = [ops.to_pf(t.to(dev)) for t in xs]
    back = ops.from_pf(pfs[0])
    ok = check("roundtrip %s" % (tuple(xs[0].shape),), back, xs[0].double(), 1e-7)
    pc = ops.pack_conv_pf(w.to(dev), split, stride, transposed, bias.to(dev))
    opf, on = ops.conv_pf(pfs, pc, act, out="both")
    torch.cuda.synchronize()
    name = "%s%dd %d->%d k%d s%d %s src%d" % ("deconv" if transposed else "conv", nd, cin, cout, k, stride, tuple(shape), nsrc)
    ok &= check(name + " nchw", on, want)
    ok &= check(name + " pf", ops.from_pf(opf), want)
    return ok


ok = True
ok &= run(32, 32, 3, 2, (24, 78))
ok &= run(32, 32, 3, 2, (96, 312))
ok &= run(32, 32, 3, 2, (48, 624))
ok &= run(64, 32, 3, 2, (40, 100), nsrc=2)
ok &= run(24, 24, 3, 3, (6, 12, 40))
ok &= run(48, 24, 1, 3, (6, 12, 40), nsrc=2)
ok &= run(16, 40, 3, 2, (30, 50), act=None, B=2)
ok &= run(32, 32, 3, 2, (40, 100), stride=2)
ok &= run(32, 16, 4, 2, (20, 50), stride=2, transposed=True)
ok &= run(24, 8, 4, 3, (5, 10, 20), stride=2, transposed=True)
ok &= run(96, 96, 3, 2, (24, 40))
print("ALL OK" if ok else "SOME FAILED")
