import sys, torch
sys.path.insert(0, ".")
from esmstereo_b200 import layers
torch.manual_seed(0)
lay = layers.SMLayer(16, 7, 2).cuda().eval()
x = torch.randn(1, 16, 96, 312, device="cuda")
for _ in range(3):
    y = lay(x, extra_residual=x)
torch.cuda.synchronize()
