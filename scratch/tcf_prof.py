import sys, torch
sys.path.insert(0, ".")
from esmstereo_b200 import ops
dev = "cuda"
cin, cout, k, nd, shape = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), tuple(int(v) for v in sys.argv[5].split("x"))
w = torch.randn(cout, cin, *((k,) * nd), device=dev) * 0.05
x = torch.randn(1, cin, *shape, device=dev)
pc = ops.pack_conv_pf(w, [cin], 1, False, None, None)
pf = ops.to_pf(x)
for _ in range(2):
    ops.conv_pf(pf, pc, "gelu", out=sys.argv[6] if len(sys.argv) > 6 else "pf")
    torch.cuda.synchronize()
