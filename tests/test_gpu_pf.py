"""GPU parity of the PF layout and the flat tcgen05 engine (conv_tcf.cu) through the C ABI: esm_pf_from_nchw_f32 /
esm_pf_to_nchw_f32 / esm_pack_conv_weight_pf_f32 / esm_conv_pf_f32 against a float64 torch convolution of the same
layer (BasicConv, submodule.py:12-38; ConvTranspose k4 s2 p1 and the stride-2 / concatenated layers of `aggregation`
and `up_refinement`, ESMStereo.py:129-239).  Gate: 2e-5 of the tensor's max, the split-TF32 engines' gate."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from esmstereo_b200 import ops
    return ops


def _rel(got, want):
    return (got.double().cpu() - want).abs().max().item() / max(want.abs().max().item(), 1e-9)


CASES = [
    # name, cin, cout, k, nd, shape, act, stride, transposed, nsrc, batch
    ("k3_2d_contiguous", 32, 32, 3, 2, (24, 78), "gelu", 1, False, 1, 1),
    ("k3_2d_band", 32, 32, 3, 2, (20, 624), "gelu", 1, False, 1, 1),
    ("k3_2d_two_sources", 64, 32, 3, 2, (40, 100), "gelu", 1, False, 2, 1),
    ("k3_3d", 24, 24, 3, 3, (6, 12, 40), "gelu", 1, False, 1, 1),
    ("k3_3d_wide_partials", 72, 72, 3, 3, (6, 12, 39), "gelu", 1, False, 1, 1),
    ("k1_3d_two_sources", 48, 24, 1, 3, (6, 12, 40), "gelu", 1, False, 2, 1),
    ("k3_2d_batch2_cout40", 16, 40, 3, 2, (30, 50), None, 1, False, 1, 2),
    ("k3_2d_padded_channels", 12, 20, 3, 2, (17, 33), "relu", 1, False, 1, 1),
    ("k3_2d_stride2", 32, 32, 3, 2, (40, 100), "gelu", 2, False, 1, 1),
    ("k3_3d_stride2_odd", 8, 24, 3, 3, (7, 11, 29), "gelu", 2, False, 1, 1),
    ("deconv_2d", 32, 16, 4, 2, (20, 50), "gelu", 2, True, 1, 1),
    ("deconv_3d", 24, 8, 4, 3, (5, 10, 20), "gelu", 2, True, 1, 1),
    ("k3_2d_two_channel_tiles", 96, 96, 3, 2, (24, 40), "silu", 1, False, 1, 1),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_pf_vs_float64(case):
    name, cin, cout, k, nd, shape, act, stride, transposed, nsrc, B = case
    ops = _ops()
    g = torch.Generator().manual_seed(len(name))
    ks = (k,) * nd
    w = torch.randn(*(((cin, cout) if transposed else (cout, cin)) + ks), generator=g) * (0.5 / (cin * k ** nd) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(B, cin, *shape, generator=g)
    conv = {2: (F.conv_transpose2d if transposed else F.conv2d), 3: (F.conv_transpose3d if transposed else F.conv3d)}[nd]
    want = conv(x.double(), w.double(), bias.double(), stride=stride, padding=(1 if (transposed or k == 3) else 0))
    want = {"gelu": F.gelu, "relu": F.relu, "silu": F.silu, None: lambda t: t}[act](want)
    split = [cin] if nsrc == 1 else [cin // 2 // 8 * 8, cin - cin // 2 // 8 * 8]
    xs = torch.split(x, split, 1)
    pfs = [ops.to_pf(t.cuda()) for t in xs]
    assert torch.equal(ops.from_pf(pfs[0]).cpu(), xs[0]), "PF round trip must be exact (hi + lo == x)"
    pc = ops.pack_conv_pf(w.cuda(), split, stride, transposed, bias.cuda())
    n0 = ops.lib().esm_tcf_conv_launches()
    opf, on = ops.conv_pf(pfs, pc, act, out="both")
    assert ops.lib().esm_tcf_conv_launches() == n0 + 1
    assert tuple(on.shape) == tuple(want.shape)
    assert _rel(on, want) < 2e-5, name
    back = ops.from_pf(opf)
    assert torch.equal(back, on.contiguous()), "PF and NCHW outputs of one launch must agree bit for bit"
    # the PF output is a valid PF tensor: its border is zero, so it can feed the next layer directly
    if not transposed and stride == 1 and k == 3 and nsrc == 1 and cin == cout:
        o2 = ops.conv_pf([opf], pc, act, out="nchw")
        want2 = {"gelu": F.gelu, "relu": F.relu, "silu": F.silu, None: lambda t: t}[act](
            conv(want, w.double(), bias.double(), stride=1, padding=1))
        assert _rel(o2, want2) < 4e-5, name + " (chained)"


def test_conv_pf_residual_and_crop():
    """Transposed layer cropped to its skip tensor (ESMStereo.py:172,230) and a residual add on both output formats."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    w = torch.randn(24, 16, 4, 4, 4, generator=g) * 0.05
    x = torch.randn(1, 24, 2, 8, 12, generator=g)
    want = F.gelu(F.conv_transpose3d(x.double(), w.double(), None, stride=2, padding=1))[:, :, :3, :16, :23]
    pc = ops.pack_conv_pf(w.cuda(), [24], 2, True, None)
    got = ops.conv_pf([ops.to_pf(x.cuda())], pc, "gelu", out="nchw", out_size=(3, 16, 23))
    assert tuple(got.shape) == (1, 16, 3, 16, 23) and _rel(got, want) < 2e-5
    w2 = torch.randn(16, 16, 3, 3, generator=g) * 0.08
    x2 = torch.randn(2, 16, 21, 37, generator=g)
    want2 = F.conv2d(x2.double(), w2.double(), None, padding=1) + x2.double()
    pc2 = ops.pack_conv_pf(w2.cuda(), [16], 1, False, None)
    pf2 = ops.to_pf(x2.cuda())
    got_n = ops.conv_pf([pf2], pc2, None, out="nchw", residual=x2.cuda())
    got_p = ops.from_pf(ops.conv_pf([pf2], pc2, None, out="pf", residual=pf2))
    assert _rel(got_n, want2) < 2e-5 and _rel(got_p, want2) < 2e-5


def test_static_rule_routes_wide_layers_to_tcf():
    """ops.conv hands wide k3 layers to the flat engine by a static rule (same layer -> same engine, every run)."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    w = torch.randn(40, 40, 3, 3, 3, generator=g) * 0.03
    x = torch.randn(1, 40, 6, 12, 39, generator=g)
    pc = ops.pack_conv(w.cuda(), 1, 1, False, None, None)
    n0 = ops.lib().esm_tcf_conv_launches()
    got = ops.conv(x.cuda(), pc, "gelu")
    assert ops.lib().esm_tcf_conv_launches() == n0 + 1
    want = F.gelu(F.conv3d(x.double(), w.double(), None, padding=1))
    assert _rel(got, want) < 2e-5
    got32 = ops.conv(x.cuda(), pc, "gelu", fp32_only=True)  # pinned layers stay on the FP32 pipe
    assert ops.lib().esm_tcf_conv_launches() == n0 + 1 and _rel(got32, want) < 2e-5
