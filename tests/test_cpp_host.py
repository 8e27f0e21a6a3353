"""The C++ host (host/esm_host.cpp) driving libesm_b200.so with no Python / torch in the process: an engine exported
from the Python model (esmstereo_b200/engine.py) replayed from C++ must reproduce the Python forward bit for bit, and the
ROS node's pre / post-processing around it (kitti_publisher_cuda_node.cpp:136-175,385-404) must match its restatement."""
import contextlib
import ctypes as C
import io
import json
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "esm_host")


def test_host_binary_is_built_and_links_the_c_abi():
    if not os.path.exists(HOST):
        subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True)
    out = subprocess.run(["ldd", HOST], capture_output=True, text=True).stdout
    assert "libesm_b200.so" in out and "libcudart" in out
    assert "libtorch" not in out and "libpython" not in out and "libc10" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("name,backbone,cv", [("ESMStereo", "efficientnet_b2", 4), ("ESMStereo_confidence", "mobilenetv2_100", 16)])
def test_cpp_host_replays_the_python_forward(tmp_path, name, backbone, cv):
    from esmstereo_b200 import __models__, ops
    from esmstereo_b200.engine import export_engine
    from esmstereo_b200.weights import fill_deterministic, synthetic_pair
    from oracle.esm_oracle import EsmOracle
    if not os.path.exists(HOST):
        subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True)
    Hp, Wp, h, w = 128, 256, 101, 230
    with contextlib.redirect_stdout(io.StringIO()):
        m = __models__[name](192, True, False, backbone, cv)
    sd = fill_deterministic(m.state_dict(), seed=0)
    conf = name == "ESMStereo_confidence"
    orc = EsmOracle(sd, 192, True, False, backbone, cv, confidence=conf)
    sd = orc.calibrate(*synthetic_pair(1, Hp, Wp, shift=7, seed=0))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    kw = {} if conf else {"train_status": False}
    eng = str(tmp_path / "model.esmeng")
    info = export_engine(m, (1, 3, Hp, Wp), eng, **kw)
    assert info["calls"] > 50 and os.path.getsize(eng) == info["file_bytes"]
    g = torch.Generator().manual_seed(3)
    base = (torch.rand(h, w + 9, 3, generator=g) * 255).to(torch.uint8)
    left, right = base[:, 9:].contiguous(), base[:, :w].contiguous()  # a 9-pixel shift
    (tmp_path / "l.u8").write_bytes(left.numpy().tobytes())
    (tmp_path / "r.u8").write_bytes(right.numpy().tobytes())
    run = subprocess.run([HOST, eng, str(tmp_path / "l.u8"), str(tmp_path / "r.u8"), str(h), str(w), str(tmp_path / "out"), "3"],
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    stats = json.loads(run.stdout.strip().splitlines()[-1])
    assert stats["network"] == [Hp, Wp] and stats["calls_per_frame"] == info["calls"]
    got = torch.from_numpy(np.fromfile(str(tmp_path / "out.disp.f32"), dtype=np.float32).reshape(Hp, Wp).copy())
    got16 = torch.from_numpy(np.fromfile(str(tmp_path / "out.u16"), dtype=np.uint16).reshape(h, w).astype(np.int32))
    # the same frame through Python: device pre-processing as the node pads (bottom / right, black pixels normalised)
    L = ops.lib()
    mean, std = (C.c_float * 3)(0.485, 0.456, 0.406), (C.c_float * 3)(0.229, 0.224, 0.225)
    ins = []
    for img in (left, right):
        x = torch.empty(1, 3, Hp, Wp, device="cuda")
        ops.check(L.esm_preprocess_u8_f32(img.cuda().data_ptr(), x.data_ptr(), 1, h, w, Hp, Wp, 0, 0, 1, mean, std, None), "preprocess")
        ins.append(x)
    with torch.no_grad():
        out = m(*ins, **kw)
    want = (out[0] if conf else out[-1])[0]
    assert torch.equal(got, want.cpu()), "C++ replay differs from the Python forward: max |d| = %g" % float((got - want.cpu()).abs().max())
    want16 = ops.disparity_publish_u16(want.contiguous(), (h, w), 192.0, 256.0).cpu().to(torch.int32)
    assert torch.equal(got16, want16)
