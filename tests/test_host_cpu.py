"""CPU-side checks: the C-ABI library loads and exports every symbol include/esm_b200.h declares,
argument validation works without a GPU, the module tree is key-compatible with the reference."""
import contextlib
import io
import os
import re

import pytest
import torch

from tests.helpers import GOLDEN_NAMES, golden_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from esmstereo_b200 import _lib
    handle = _lib.lib()
    header = open(os.path.join(ROOT, "include", "esm_b200.h")).read()
    declared = set(re.findall(r"\b(esm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(handle, name), "libesm_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert handle.esm_version() >= 100


def test_shipped_library_sass_has_tcgen05_and_no_abi_printf():
    """The product library is sm_100a code that really uses the Blackwell paths (tcgen05.mma = UTCHMMA, tensor-memory
    loads / stores, TMA bulk copies), and no kernel carries a device printf: a vprintf ABI call inside the mbarrier wait
    loops cost the whole forward 1.9 % (-DTC_DEADLOCK_PRINTF builds are for debugging only)."""
    import shutil
    import subprocess
    from esmstereo_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    if os.environ.get("ESM_LIB"):
        pytest.skip("an A/B build is loaded")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for op in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "NANOSLEEP.SYNCS"):
        assert op in sass, "no %s in the shipped library" % op
    assert "vprintf" not in sass, "a device printf (ABI call) is compiled into the shipped kernels"
    assert "CALL.ABS" not in sass, "an ABI call is compiled into the shipped kernels"


def test_argument_errors_need_no_gpu():
    from esmstereo_b200 import _lib
    L = _lib.lib()
    assert L.esm_gwc_volume_f32(None, None, None, 1, 64, 4, 8, 4, 32, None) == -1
    assert b"null" in L.esm_last_error()
    assert L.esm_conv_f32(None, None) == -1
    # fp32 pack [phase][tap][Cin_pad][Cout_pad] + (Cin >= 8) the split-TF32 slabs of the streamed tcgen05 engine:
    # phases * taps * ceil(Cin/8) slabs of [hi|lo][2][round_up(Cout,8)][4]
    n = L.esm_packed_weight_elems(8, 32, 3, 3, 3, 0)
    # ... + (stride-1 k1 / k3 layers) the resident-weight image of conv_tc.cu, 32-float aligned:
    # [cot][cg][kd][hi|lo][2][taps * COT][4] with COT = 8 here
    assert n == 27 * 32 * 8 + 27 * 4 * 16 * 8 + 1 * 4 * 3 * 2 * (9 * 8) * 8
    # 2D k3 with Cout % 32 == 0 and Cin <= 64: the image of the kh-in-K kernel, [cot][cg][kh][hi|lo][2][96][4]
    assert L.esm_packed_weight_elems(32, 32, 1, 3, 3, 0) == 9 * 32 * 32 + 9 * 4 * 16 * 32 + 1 * 4 * 3 * 2 * 96 * 8
    # transposed layers carry no image
    assert L.esm_packed_weight_elems(24, 40, 4, 4, 4, 1) == 8 * 8 * 40 * 24 + 8 * 8 * 5 * 16 * 24
    assert L.esm_packed_weight_elems(40, 72, 4, 4, 4, 1) == 8 * 8 * 72 * 40 + 8 * 8 * 9 * 16 * 40
    assert L.esm_packed_weight_elems(72, 8, 1, 1, 1, 0) == 8 * 80 + 1 * 1 * 16 * 72   # Cout 72 -> 2 CTAs x 40 channels
    assert L.esm_packed_weight_elems(1, 24, 4, 4, 4, 1) == 8 * 8 * 24 * 4 + 8 * 8 * 3 * 16 * 8   # Cout=1 padded to 4 (16-byte weight rows)
    assert L.esm_packed_weight_elems(32, 1, 1, 5, 5, 0) == 25 * 1 * 32      # Cin=1 keeps CK=1, no tensor-core slabs


def test_no_cpu_fallback():
    from esmstereo_b200 import ops
    with pytest.raises(TypeError):
        ops.regression_top2(torch.zeros(1, 4, 2, 2))
    if not torch.cuda.is_available():
        from esmstereo_b200 import _lib
        assert _lib.lib().esm_device_info(None, None, None) == -2  # ESM_ERR_CUDA, loudly


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_state_dict_keys_match_reference(name):
    from esmstereo_b200 import __models__
    cfg = golden_config(name)
    with contextlib.redirect_stdout(io.StringIO()):
        m = __models__[cfg["model"]](192, cfg["gwc"], cfg["norm_correlation"], cfg["backbone"], cfg["cv_scale"])
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == cfg["keys"]


def test_constructor_contract():
    from esmstereo_b200 import __models__
    assert set(__models__) == {"ESMStereo", "ESMStereo_trt", "ESMStereo_confidence"}
    with pytest.raises(NameError):  # the reference's `pirnt` typo path, ESMStereo.py:599
        __models__["ESMStereo"](192, True, False, "efficientnet_b2", 5)


def test_tc_gelu_polynomial():
    """The tcgen05 epilogues evaluate GELU as x - h / h with h = 0.5 x erfc(|x|/sqrt 2), erfc(t) = 2^p(t), p a degree-8
    fit of log2(erfcx(t)) - t^2 log2(e) on [0, 4] (csrc/tc_common.cuh: tc_gelu).  Replays the fp32 Horner evaluation
    with the coefficients parsed from the CUDA source and bounds the error against torch's exact-erf GELU."""
    import os
    import re
    import numpy as np
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "esmstereo_b200", "csrc", "tc_common.cuh")).read()
    body = src[src.index("float tc_gelu(float x)"):]
    body = body[:body.index("ex2.approx")]
    coef = [np.float32(c) for c in re.findall(r"(-?\d\.\d+e[+-]\d+)f", body)]
    assert len(coef) == 9
    x32 = torch.linspace(-8, 8, 400001, dtype=torch.float64).to(torch.float32).numpy()
    x = torch.from_numpy(x32.astype(np.float64))  # the same (fp32-representable) inputs for the exact reference
    t = np.minimum(np.abs(x32) * np.float32(0.70710678118654752440), np.float32(4.0)).astype(np.float32)
    q = np.full_like(t, coef[0])
    for c in coef[1:]:
        q = (q * t + c).astype(np.float32)
    e = np.exp2(q.astype(np.float64)).astype(np.float32)
    h = (np.float32(0.5) * x32) * e
    got = np.where(x32 >= 0, x32 - h, h).astype(np.float64)
    want = torch.nn.functional.gelu(x).numpy()
    small = np.abs(x32) <= 4
    assert np.abs(got - want)[small].max() < 1.5e-7                       # an ulp of an O(1) activation
    assert (np.abs(got - want)[~small] / np.abs(x32[~small])).max() < 6e-8  # beyond: half an ulp of x
    pos = x32 > 1e-3
    assert (np.abs(got - want)[pos] / want[pos]).max() < 5e-7
