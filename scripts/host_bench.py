"""Export the KITTI-shaped engine and time the C++ host on it (frames/s end to end: H2D of two uint8 images, device
pre-processing, graph replay, median + 16U post-processing, D2H), the B200-native run of the reference's ROS publisher
loop (kitti_publisher_cuda_node.cpp:323-430).   python scripts/host_bench.py [H W h w]"""
import contextlib, io, json, os, subprocess, sys, tempfile
os.environ.setdefault("ESM_BACKBONE", "standin")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esmstereo_b200 import __models__
from esmstereo_b200.engine import export_engine
from esmstereo_b200.weights import fill_deterministic

Hp, Wp, h, w = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (384, 1248, 375, 1242)
with contextlib.redirect_stdout(io.StringIO()):
    m = __models__["ESMStereo"](192, True, False, "efficientnet_b2", 4)
m.load_state_dict(fill_deterministic(m.state_dict()))
m = m.cuda().eval()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with tempfile.TemporaryDirectory() as td:
    eng = os.path.join(td, "kitti.esmeng")
    info = export_engine(m, (1, 3, Hp, Wp), eng, train_status=False)
    g = torch.Generator().manual_seed(0)
    img = (torch.rand(h, w + 23, 3, generator=g) * 255).to(torch.uint8)
    open(os.path.join(td, "l.u8"), "wb").write(img[:, 23:].contiguous().numpy().tobytes())
    open(os.path.join(td, "r.u8"), "wb").write(img[:, :w].contiguous().numpy().tobytes())
    del m
    torch.cuda.empty_cache()
    run = subprocess.run([os.path.join(root, "host", "esm_host"), eng, os.path.join(td, "l.u8"), os.path.join(td, "r.u8"), str(h), str(w),
                          os.path.join(td, "out"), "200"], capture_output=True, text=True)
    sys.stderr.write(run.stderr)
    stats = json.loads(run.stdout.strip().splitlines()[-1])
    stats["engine"] = {k: info[k] for k in ("calls", "segments", "reserved_bytes", "state_bytes", "file_bytes")}
    print(json.dumps(stats))
