"""GPU parity of every C-ABI operator against a torch-CPU fp32 statement of the same op
(the oracle's primitives where they exist, `torch.nn.functional` for the convolutions).

Tolerances (fp32 kernels, different summation order than MKL-DNN): conv outputs 2e-5 relative to the
tensor's max magnitude; cost volume 1e-6 (un-contracted arithmetic, expected bit-equal for cpg=2);
top-2 indices bit-exact.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.esm_oracle import EsmOracle

pytestmark = pytest.mark.gpu

ACTS = {None: lambda x: x, "gelu": F.gelu, "relu": F.relu, "silu": F.silu, "sigmoid": torch.sigmoid,
        "2sigmoid": lambda x: 2 * torch.sigmoid(x), "relu6": F.relu6}


def _ops():
    from esmstereo_b200 import ops
    return ops


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape) * 7919)
    return torch.randn(*shape, generator=g) * scale


def make_layer(cin, cout, k, nd, transposed=False, bn=True, bias=False, seed=0):
    ks = (k,) * nd if isinstance(k, int) else tuple(k)
    wshape = ((cin, cout) if transposed else (cout, cin)) + ks
    fan = cin * int(np.prod(ks))
    w = rnd(*wshape, seed=seed) * (2.0 / fan) ** 0.5
    p = {"w": w, "bias": rnd(cout, seed=seed + 1) * 0.1 if bias else None, "bn": None}
    if bn:
        p["bn"] = (0.5 + torch.rand(cout), rnd(cout, seed=seed + 2) * 0.1, rnd(cout, seed=seed + 3) * 0.1,
                   0.5 + torch.rand(cout), 1e-5)
    return p


def ref_conv(x, p, stride, pad, transposed, act, nd):
    fn = {(2, False): F.conv2d, (2, True): F.conv_transpose2d, (3, False): F.conv3d, (3, True): F.conv_transpose3d}[(nd, transposed)]
    y = fn(x, p["w"], p["bias"], stride=stride, padding=pad)
    if p["bn"] is not None:
        g, b, m, v, eps = p["bn"]
        y = F.batch_norm(y, m, v, g, b, False, 0.0, eps)
    return ACTS[act](y)


def gpu_pack(p, stride, pad, transposed):
    ops = _ops()
    bn = None if p["bn"] is None else tuple(t.cuda() for t in p["bn"][:4]) + (p["bn"][4],)
    return ops.pack_conv(p["w"].cuda(), stride, pad, transposed, None if p["bias"] is None else p["bias"].cuda(), bn)


CONV_CASES = [
    # name, nd, cin, cout, k, stride, pad, transposed, act, bn, bias, in_shape (spatial), batch
    ("stem3d_32_8", 3, 32, 8, 3, 1, 1, False, "gelu", True, False, (6, 10, 36), 1),
    ("agg3d_8_8_b2", 3, 8, 8, 3, 1, 1, False, "gelu", True, False, (5, 9, 20), 2),
    ("down3d_8_24_s2", 3, 8, 24, 3, 2, 1, False, "gelu", True, False, (12, 18, 40), 1),
    ("down3d_odd_40_72_s2", 3, 40, 72, 3, 2, 1, False, "gelu", True, False, (3, 5, 7), 1),
    ("conv3d_24_24", 3, 24, 24, 3, 1, 1, False, "gelu", True, False, (6, 9, 39), 1),
    ("conv3d_72_72", 3, 72, 72, 3, 1, 1, False, "gelu", True, False, (2, 3, 10), 1),
    ("conv3d_12_12_pad", 3, 12, 12, 3, 1, 1, False, "gelu", True, False, (6, 7, 11), 1),
    ("corr_stem_1_8", 3, 1, 8, 3, 1, 1, False, "gelu", True, False, (12, 6, 10), 1),
    ("k1_3d_48_24", 3, 48, 24, 1, 1, 0, False, "gelu", True, False, (4, 6, 18), 1),
    ("deconv3d_72_40", 3, 72, 40, 4, 2, 1, True, "gelu", True, False, (2, 3, 5), 1),
    ("deconv3d_24_1", 3, 24, 1, 4, 2, 1, True, None, False, False, (4, 6, 13), 2),
    ("dm0_k5p1_1_32", 2, 1, 32, 5, 1, 1, False, "gelu", True, False, (12, 40), 1),
    ("dm3_k1p1_32_32", 2, 32, 32, 1, 1, 1, False, "gelu", True, False, (10, 38), 1),
    ("ref_s2_1_32", 2, 1, 32, 3, 2, 1, False, "gelu", True, False, (24, 80), 1),
    ("ref_s2_32_32", 2, 32, 32, 3, 2, 1, False, "gelu", True, False, (12, 40), 2),
    ("conv2d_16_16", 2, 16, 16, 3, 1, 1, False, "gelu", True, False, (9, 21), 1),
    ("deconv2d_32_32", 2, 32, 32, 4, 2, 1, True, "gelu", True, False, (6, 20), 1),
    ("deconv2d_16_1", 2, 16, 1, 4, 2, 1, True, "gelu", True, False, (6, 10), 1),
    ("deconv2d_32_1_plain", 2, 32, 1, 4, 2, 1, True, None, False, False, (12, 40), 1),
    ("tail_16_1_bias", 2, 16, 1, 3, 1, 1, False, None, False, True, (24, 80), 1),
    ("fm_16_32_silu_bias", 2, 16, 32, 3, 1, 1, False, "silu", False, True, (12, 40), 1),
    ("laf_7_16_relu", 2, 7, 16, 3, 1, 1, False, "relu", True, True, (6, 10), 1),
    ("laf_64_16_relu", 2, 64, 16, 3, 1, 1, False, "relu", True, True, (6, 10), 1),
    ("laf_16_1_2sig", 2, 16, 1, 1, 1, 0, False, "2sigmoid", True, True, (6, 10), 1),
    ("wide_row", 2, 32, 32, 3, 1, 1, False, "gelu", True, False, (4, 312), 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_matches_torch(case):
    name, nd, cin, cout, k, stride, pad, transposed, act, bn, bias, sp, B = case
    ops = _ops()
    p = make_layer(cin, cout, k, nd, transposed, bn, bias, seed=len(name))
    x = rnd(B, cin, *sp, seed=3)
    want = ref_conv(x, p, stride, pad, transposed, act, nd)
    got = ops.conv(x.cuda(), gpu_pack(p, stride, pad, transposed), act)
    assert rel(got, want) < 2e-5, name


def test_conv_concat_sources_and_crop():
    """torch.cat((a[cropped], b, c), 1) -> k1 conv (aggregation.agg_0, up_refinement.agg_0)."""
    ops = _ops()
    a_full = rnd(1, 40, 4, 6, 10, seed=1)  # deconv output before crop-to-skip
    a = a_full[:, :, :3, :5, :9]
    b = rnd(1, 40, 3, 5, 9, seed=2)
    p = make_layer(80, 40, 1, 3, seed=5)
    want = ref_conv(torch.cat((a, b), 1), p, 1, 0, False, "gelu", 3)
    got = ops.conv([a_full.cuda()[:, :, :3, :5, :9], b.cuda()], gpu_pack(p, 1, 0, False), "gelu")
    assert rel(got, want) < 2e-5
    srcs = [rnd(2, 32, 6, 20, seed=7), rnd(2, 32, 6, 20, seed=8), rnd(2, 96, 6, 20, seed=9)]
    p = make_layer(160, 32, 1, 2, seed=6)
    want = ref_conv(torch.cat(srcs, 1), p, 1, 0, False, "gelu", 2)
    got = ops.conv([s.cuda() for s in srcs], gpu_pack(p, 1, 0, False), "gelu")
    assert rel(got, want) < 2e-5
    srcs = [rnd(1, 16, 6, 10, seed=7), rnd(1, 1, 6, 10, seed=8)]  # LAFNet fusion_conv1: cat(feat, out)
    p = make_layer(17, 16, 3, 2, bias=True, seed=6)
    want = ref_conv(torch.cat(srcs, 1), p, 1, 1, False, "relu", 2)
    got = ops.conv([s.cuda() for s in srcs], gpu_pack(p, 1, 1, False), "relu")
    assert rel(got, want) < 2e-5


def test_deconv_crop_to_skip():
    ops = _ops()
    x = rnd(1, 24, 2, 8, 12, seed=1)
    p = make_layer(24, 16, 4, 3, transposed=True, seed=2)
    full = ref_conv(x, p, 2, 1, True, "gelu", 3)
    want = full[:, :, :3, :16, :23]
    got = ops.conv(x.cuda(), gpu_pack(p, 2, 1, True), "gelu", out_size=(3, 16, 23))
    assert rel(got, want) < 2e-5


def test_conv_epilogue_fusions():
    ops = _ops()
    x = rnd(2, 16, 6, 10, seed=1)
    p = make_layer(16, 16, 1, 2, bn=False, bias=True, seed=2)
    res = rnd(2, 16, 6, 10, seed=3)
    want = ref_conv(x, p, 1, 0, False, None, 2) + res
    got = ops.conv(x.cuda(), gpu_pack(p, 1, 0, False), None, residual=res.cuda())
    assert rel(got, want) < 2e-5
    # residual + second activation + scale (conf_upsample tail: sigmoid(conv1_up(x) + conf1))
    p = make_layer(16, 1, 4, 2, transposed=True, seed=4)
    res = rnd(2, 1, 12, 20, seed=5)
    want = torch.sigmoid(ref_conv(x, p, 2, 1, True, "gelu", 2) + res) * 3.0
    got = ops.conv(x.cuda(), gpu_pack(p, 2, 1, True), "gelu", residual=res.cuda(), act2="sigmoid", out_scale=3.0)
    assert rel(got, want) < 2e-5
    # broadcast multipliers over D (att, ESMStereo.py:703,711)
    v = rnd(1, 32, 4, 6, 10, seed=6)
    att_in = rnd(1, 32, 6, 10, seed=7)
    att_out = rnd(1, 8, 6, 10, seed=8)
    p = make_layer(32, 8, 3, 3, seed=9)
    want = ref_conv(v * att_in.unsqueeze(2), p, 1, 1, False, "gelu", 3) * att_out.unsqueeze(2)
    got = ops.conv(v.cuda(), gpu_pack(p, 1, 1, False), "gelu", in_mul=att_in.cuda(), out_mul=att_out.cuda())
    assert rel(got, want) < 2e-5


@pytest.mark.parametrize("r,n", [(2, 16), (4, 8)])
def test_conv_pixel_shuffle_silu(r, n):
    ops = _ops()
    x = rnd(1, n, 6, 10, seed=1)
    p = make_layer(n, n * r * r, 1, 2, bn=False, bias=True, seed=2)
    want = F.silu(F.pixel_shuffle(ref_conv(x, p, 1, 0, False, None, 2), r))
    got = ops.conv(x.cuda(), gpu_pack(p, 1, 0, False), "silu", pixel_shuffle=r)
    assert rel(got, want) < 2e-5


@pytest.mark.parametrize("B,C,H,W,D,G", [(1, 64, 6, 40, 12, 32), (2, 64, 5, 10, 12, 32), (1, 64, 4, 39, 48, 32),
                                         (1, 24, 3, 16, 5, 8)])
def test_gwc_volume(B, C, H, W, D, G):
    ops = _ops()
    L, R = rnd(B, C, H, W, seed=1), rnd(B, C, H, W, seed=2)
    want = EsmOracle({}, 192).gwc_volume(L, R, D, G)
    got = ops.build_gwc_volume(L.cuda(), R.cuda(), D, G)
    assert rel(got, want) < 1e-6
    if C // G == 2:
        assert torch.equal(got.cpu(), want), "cpg=2 volume must be bit-identical to the reference arithmetic"


def test_gwc_fused_into_stem_conv():
    ops = _ops()
    L, R = rnd(1, 64, 6, 40, seed=1), rnd(1, 64, 6, 40, seed=2)
    D = 12
    p = make_layer(32, 8, 3, 3, seed=3)
    vol = EsmOracle({}, 192).gwc_volume(L, R, D, 32)
    want = ref_conv(vol, p, 1, 1, False, "gelu", 3)
    pc = gpu_pack(p, 1, 1, False)
    got = ops.conv([L.cuda(), R.cuda()], pc, "gelu", gwc_disp=D)
    assert rel(got, want) < 2e-5
    att = rnd(1, 32, 6, 40, seed=4)
    want = ref_conv(vol * att.unsqueeze(2), p, 1, 1, False, "gelu", 3)
    got = ops.conv([L.cuda(), R.cuda()], pc, "gelu", gwc_disp=D, in_mul=att.cuda())
    assert rel(got, want) < 2e-5


@pytest.mark.parametrize("B,C,H,W,D", [(1, 64, 6, 40, 12), (2, 64, 3, 10, 12)])
def test_norm_corr_volume(B, C, H, W, D):
    ops = _ops()
    L, R = rnd(B, C, H, W, seed=1), rnd(B, C, H, W, seed=2)
    want = EsmOracle({}, 192).norm_corr_volume(L, R, D)
    got = ops.build_norm_correlation_volume(L.cuda(), R.cuda(), D)
    assert rel(got, want) < 1e-4  # the BASELINE.json volume gate


def test_regression_top2_indices_bit_exact():
    ops = _ops()
    cost = rnd(2, 48, 24, 78, seed=1)
    cost[0, :, 0, 0] = 0.0           # all ties
    cost[0, 5, 0, 1] = cost[0, 9, 0, 1] = 7.0   # tie for the maximum
    cost[0, 3, 0, 2] = 9.0
    cost[0, 11, 0, 2] = cost[0, 20, 0, 2] = 8.0  # tie for second
    want, widx = EsmOracle.regression_top2(cost)
    got, gidx = ops.regression_top2(cost.cuda(), return_indices=True)
    assert torch.equal(gidx.cpu().long(), widx), "top-2 indices must be bit-exact"
    assert float((got.cpu() - want).abs().max()) < 1e-5
    assert torch.equal(ops.regression_topk(cost.cuda(), None, 2), got)


def test_disparity_regression_no_softmax():
    ops = _ops()
    cost = rnd(2, 12, 6, 10, seed=1)
    want = EsmOracle.disparity_regression(cost).squeeze(1)
    got = ops.disparity_regression(cost.cuda(), 12)
    assert got.shape == want.shape and rel(got, want) < 1e-6


@pytest.mark.parametrize("f", [2, 4])
def test_bilinear_add(f):
    ops = _ops()
    prev, res = rnd(2, 1, 6, 10, seed=1), rnd(2, 1, 6 * f, 10 * f, seed=2)
    want = (F.interpolate(prev, scale_factor=f, mode="bilinear", align_corners=False) + res) * 4
    got = ops.bilinear_add(prev.cuda(), res.cuda(), f, 4.0)
    assert rel(got, want) < 1e-6


@pytest.mark.parametrize("hw", [(6, 10), (24, 78), (13, 21)])
def test_final_assembly_fused_into_conv1_up(hw):
    """`F.interpolate(prev, x2, bilinear) + conv1_up(x)` and the final `* 4` (ESMStereo.py:307,316,745) in the epilogue of
    conv1_up's sub-pixel form (pixel_shuffle = 2 with a low-resolution residual): bit-identical to the conv followed by
    esm_bilinear_add_f32, and equal to the torch statement."""
    from esmstereo_b200 import layers
    ops = _ops()
    h, w = hw
    torch.manual_seed(h)
    up = layers.BasicConv(32, 1, deconv=True, is_3d=False, bn=False, gelu=False, kernel_size=4, padding=1, stride=2)
    x, prev = rnd(2, 32, h, w, seed=3), rnd(2, 1, h, w, seed=4) * 10
    want = (F.interpolate(prev, scale_factor=2, mode="bilinear", align_corners=False)
            + F.conv_transpose2d(x, up.conv.weight.detach(), None, stride=2, padding=1)) * 4
    up = up.cuda().eval()
    fused = up(x.cuda(), bilinear_prev=prev.cuda(), final_scale=4.0)
    two = ops.bilinear_add(prev.cuda(), up(x.cuda()), 2, 4.0)
    assert torch.equal(fused, two)
    assert rel(fused, want) < 2e-5


@pytest.mark.parametrize("C", [8, 16])
def test_shufflemixer_block(C):
    """FMBlock (2 SMLayers + conv tail) against the oracle's restatement of shufflemixer.py."""
    from esmstereo_b200 import layers
    torch.manual_seed(C)
    blk = layers.FMBlock(C, 7, 2)
    sd = {k: (torch.randn_like(v) * 0.3 + (1.0 if k.endswith("body.weight") else 0.0)) for k, v in blk.state_dict().items()}
    blk.load_state_dict(sd)
    x = rnd(2, C, 12, 40, seed=1)
    want = EsmOracle({"b." + k: v for k, v in sd.items()}, 192).fm_block(x, "b")
    got = blk.cuda().eval()(x.cuda())
    assert rel(got, want) < 2e-5
    # the fused SMLayer kernel (one launch) against the two half kernels: same arithmetic, bit for bit; odd sizes / partial tiles
    for shape in ((2, C, 12, 40), (1, C, 21, 45), (1, C, 96, 312)):
        xs = rnd(*shape, seed=shape[2]).cuda()
        for lay in (blk.net[0], blk.net[1]):
            lay.fused = True
            a = lay(xs, extra_residual=xs)
            lay.fused = False
            b = lay(xs, extra_residual=xs)
            lay.fused = True
            assert torch.equal(a, b), shape


def test_laf_pieces():
    ops = _ops()
    cost = rnd(2, 12, 6, 10, seed=1)
    nrm = torch.sqrt((cost ** 2).sum(1, keepdim=True) + 1e-6)
    want = torch.topk(F.softmax(-(cost / nrm) * 100, 1), k=7, dim=1).values
    assert rel(ops.laf_cost_top7(cost.cuda()), want) < 1e-5
    towers = [rnd(2, 16, 6, 10, seed=s) for s in (2, 3, 4)]
    atts = [rnd(2, 1, 6, 10, seed=s) for s in (5, 6, 7)]
    a = F.softmax(torch.cat(atts, 1), 1)
    want = torch.cat([t * a[:, i:i + 1] for i, t in enumerate(towers)], 1)
    got = ops.laf_attention(*[t.cuda() for t in towers], *[t.cuda() for t in atts])
    assert rel(got, want) < 1e-6


def test_conf_convex_up4():
    ops = _ops()
    feat, conf = rnd(2, 16, 6, 10, seed=1), rnd(2, 1, 6, 10, seed=2)
    w, b = rnd(16, 9, 4, 4, seed=3) * 0.3, rnd(9, seed=4) * 0.1
    sfm = F.softmax(F.conv_transpose2d(feat, w, b, stride=4), 1)
    nb = F.interpolate(F.unfold(conf, 3, 1, 1).reshape(2, 9, 6, 10), (24, 40), mode="nearest")
    want = (nb * sfm).sum(1, keepdim=True)
    got = ops.conf_convex_up4(feat.cuda(), conf.cuda(), w.cuda(), b.cuda())
    assert rel(got, want) < 1e-5


def test_errors_are_loud():
    ops = _ops()
    with pytest.raises(TypeError):
        ops.build_gwc_volume(torch.zeros(1, 64, 4, 8), torch.zeros(1, 64, 4, 8), 4, 32)  # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        ops.build_gwc_volume(torch.zeros(1, 30, 4, 8).cuda(), torch.zeros(1, 30, 4, 8).cuda(), 4, 32)
    p = make_layer(8, 8, 3, 2)
    pc = gpu_pack(p, 1, 1, False)
    with pytest.raises(AssertionError):
        ops.conv(torch.zeros(1, 9, 4, 8).cuda(), pc)


# ---- tensor-core (tcgen05) path of the k3 s1 p1 convolutions: forced on, every structural variant ----
TC_CASES = [
    # name, nd, cin, cout, in_shape, batch
    ("tc3d_8_8", 3, 8, 8, (7, 37, 95), 2),        # COT=8, TZ=3: partial z tile, several strips and y ranges
    ("tc3d_32_8", 3, 32, 8, (6, 10, 36), 1),
    ("tc3d_24_24", 3, 24, 24, (5, 20, 61), 1),    # COT=24, TZ=1
    ("tc3d_16_16", 3, 16, 16, (4, 9, 33), 1),     # COT=16
    ("tc3d_24_40", 3, 24, 40, (3, 6, 20), 1),     # two channel tiles of 24 (padded)
    ("tc3d_12_12_padded_cin", 3, 12, 12, (4, 7, 11), 1),
    ("tc2d_32_32", 2, 32, 32, (50, 200), 1),      # two channel tiles of 16, ragged 4-group stages
    ("tc2d_96_48", 2, 96, 48, (12, 40), 2),
    ("tc2d_16_8", 2, 16, 8, (9, 21), 1),
    ("tc2d_24_64", 2, 24, 64, (31, 64), 1),
    ("tc2d_wide", 2, 32, 32, (4, 312), 1),
]


@pytest.fixture
def tc_forced(monkeypatch):
    """Force the tensor-core path on every eligible layer (otherwise it has to win the on-device timing)."""
    monkeypatch.setenv("ESM_TC_FORCE", "1")
    monkeypatch.setenv("ESM_TC", "3")
    from esmstereo_b200 import _lib
    return _lib.lib().esm_tc_conv_launches


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv_tensor_core_path(case, tc_forced, monkeypatch):
    """Split-TF32 tcgen05 path: same gate as the FP32-pipe kernels (2e-5 of the tensor's max)."""
    name, nd, cin, cout, sp, B = case
    ops = _ops()
    p = make_layer(cin, cout, 3, nd, seed=len(name))
    x = rnd(B, cin, *sp, seed=3)
    want = ref_conv(x, p, 1, 1, False, "gelu", nd)
    pc = gpu_pack(p, 1, 1, False)
    n0 = tc_forced()
    got = ops.conv(x.cuda(), pc, "gelu")
    assert tc_forced() == n0 + 1, "layer did not take the tensor-core path"
    assert rel(got, want) < 2e-5, name
    # single-pass TF32 fast mode: operands rounded to 10 mantissa bits
    monkeypatch.setenv("ESM_TC", "1")
    got1 = ops.conv(x.cuda(), pc, "gelu")
    assert tc_forced() == n0 + 2
    assert rel(got1, want) < 3e-3, name


def test_conv_tensor_core_fusions(tc_forced):
    ops = _ops()
    # concat of two sources + residual + second activation + scale
    a, b = rnd(2, 16, 11, 45, seed=1), rnd(2, 8, 11, 45, seed=2)
    p = make_layer(24, 16, 3, 2, bias=True, seed=3)
    res = rnd(2, 16, 11, 45, seed=4)
    want = torch.sigmoid(ref_conv(torch.cat((a, b), 1), p, 1, 1, False, "gelu", 2) + res) * 3.0
    n0 = tc_forced()
    got = ops.conv([a.cuda(), b.cuda()], gpu_pack(p, 1, 1, False), "gelu", residual=res.cuda(), act2="sigmoid", out_scale=3.0)
    assert tc_forced() == n0 + 1
    assert rel(got, want) < 2e-5
    # cropped 3D view as input, out_mul broadcast over D
    full = rnd(1, 8, 6, 12, 40, seed=5)
    v = full[:, :, :5, :11, :37]
    att = rnd(1, 8, 11, 37, seed=6)
    p = make_layer(8, 8, 3, 3, seed=7)
    want = ref_conv(v, p, 1, 1, False, "gelu", 3) * att.unsqueeze(2)
    got = ops.conv(full.cuda()[:, :, :5, :11, :37], gpu_pack(p, 1, 1, False), "gelu", out_mul=att.cuda())
    assert tc_forced() == n0 + 2
    assert rel(got, want) < 2e-5


@pytest.mark.parametrize("H,W,D", [(6, 40, 12), (20, 100, 12), (5, 33, 7)])
def test_gwc_fused_tensor_core(H, W, D, tc_forced):
    ops = _ops()
    L, R = rnd(1, 64, H, W, seed=1), rnd(1, 64, H, W, seed=2)
    p = make_layer(32, 8, 3, 3, seed=3)
    vol = EsmOracle({}, 192).gwc_volume(L, R, D, 32)
    want = ref_conv(vol, p, 1, 1, False, "gelu", 3)
    n0 = tc_forced()
    got = ops.conv([L.cuda(), R.cuda()], gpu_pack(p, 1, 1, False), "gelu", gwc_disp=D)
    assert tc_forced() == n0 + 1
    assert rel(got, want) < 2e-5


@pytest.mark.parametrize("kind", ["gwc_stem", "agg3d", "conv2d_24", "k1_96_32"])
def test_tensor_core_resident_engine_is_race_free(kind, tc_forced):
    """The taps-in-N tcgen05 kernel publishes its producers' shared-memory operand tiles to the tensor core with a
    proxy fence on the CONSUMER side of the mbarrier release / acquire chain (conv_tc.cu: the writer-side fence's
    MEMBAR waited for the producers' loads in flight).  The kernel has no atomics, so any stale operand read would
    show as a launch whose output differs from the others: 40 launches at shapes with many ring wrap-arounds must be
    bit-identical, and agree with the fp32 reference."""
    ops = _ops()
    if kind == "gwc_stem":
        L, R = rnd(1, 64, 48, 156, seed=1), rnd(1, 64, 48, 156, seed=2)
        p = make_layer(32, 8, 3, 3, seed=3)
        want = ref_conv(EsmOracle({}, 192).gwc_volume(L, R, 12, 32), p, 1, 1, False, "gelu", 3)
        args, kw = ([L.cuda(), R.cuda()], gpu_pack(p, 1, 1, False), "gelu"), dict(gwc_disp=12)
    elif kind == "agg3d":
        x = rnd(1, 8, 12, 48, 156, seed=4)
        p = make_layer(8, 8, 3, 3, seed=5)
        want = ref_conv(x, p, 1, 1, False, "gelu", 3)
        args, kw = (x.cuda(), gpu_pack(p, 1, 1, False), "gelu"), {}
    elif kind == "conv2d_24":
        x = rnd(2, 24, 96, 312, seed=6)
        p = make_layer(24, 24, 3, 2, seed=7)
        want = ref_conv(x, p, 1, 1, False, "gelu", 2)
        args, kw = (x.cuda(), gpu_pack(p, 1, 1, False), "gelu"), {}
    else:
        x = rnd(1, 96, 96, 312, seed=8)
        p = make_layer(96, 32, 1, 2, seed=9)
        want = ref_conv(x, p, 1, 0, False, "gelu", 2)
        args, kw = (x.cuda(), gpu_pack(p, 1, 0, False), "gelu"), {}
    n0 = tc_forced()
    first = ops.conv(*args, **kw).clone()
    assert tc_forced() == n0 + 1
    assert rel(first, want) < 2e-5
    for i in range(40):
        again = ops.conv(*args, **kw)
        assert torch.equal(again, first), "launch %d differs from the first" % i


TC_K1_CASES = [
    # name, nd, cin, cout, in_shape, batch
    ("tck1_16_64", 2, 16, 64, (37, 100), 1),
    ("tck1_96_32", 2, 96, 32, (20, 70), 2),
    ("tck1_112_32", 2, 112, 32, (9, 33), 1),
    ("tck1_32_16", 2, 32, 16, (12, 312), 1),
    ("tck1_64_40", 2, 64, 40, (7, 19), 1),
    ("tck1_3d_48_24", 3, 48, 24, (4, 6, 18), 1),
]


@pytest.mark.parametrize("case", TC_K1_CASES, ids=[c[0] for c in TC_K1_CASES])
def test_conv_tensor_core_pointwise(case, tc_forced):
    name, nd, cin, cout, sp, B = case
    ops = _ops()
    p = make_layer(cin, cout, 1, nd, seed=len(name))
    x = rnd(B, cin, *sp, seed=3)
    want = ref_conv(x, p, 1, 0, False, "gelu", nd)
    n0 = tc_forced()
    got = ops.conv(x.cuda(), gpu_pack(p, 1, 0, False), "gelu")
    assert tc_forced() == n0 + 1, "layer did not take the tensor-core path"
    assert rel(got, want) < 2e-5, name


def test_conv_tensor_core_pointwise_concat(tc_forced):
    """aggregation.agg_0.0: cat(cropped deconv output, skip) -> k1 conv, on the tensor-core path."""
    ops = _ops()
    a_full = rnd(1, 40, 4, 6, 10, seed=1)
    a = a_full[:, :, :3, :5, :9]
    b = rnd(1, 40, 3, 5, 9, seed=2)
    p = make_layer(80, 40, 1, 3, seed=5)
    want = ref_conv(torch.cat((a, b), 1), p, 1, 0, False, "gelu", 3)
    n0 = tc_forced()
    got = ops.conv([a_full.cuda()[:, :, :3, :5, :9], b.cuda()], gpu_pack(p, 1, 0, False), "gelu")
    assert tc_forced() == n0 + 1
    assert rel(got, want) < 2e-5
    srcs = [rnd(2, 32, 6, 20, seed=7), rnd(2, 32, 6, 20, seed=8), rnd(2, 96, 6, 20, seed=9)]
    p = make_layer(160, 32, 1, 2, seed=6)
    want = ref_conv(torch.cat(srcs, 1), p, 1, 0, False, "gelu", 2)
    got = ops.conv([s.cuda() for s in srcs], gpu_pack(p, 1, 0, False), "gelu")
    assert tc_forced() == n0 + 2
    assert rel(got, want) < 2e-5


# ---- streamed-weight tcgen05 engine (conv_tcg.cu: taps in K, weights through the ring): forced on ----
TCG_CASES = [
    # name, nd, cin, cout, k, stride, pad, transposed, act, in_shape, batch
    ("tcg3d_40_40", 3, 40, 40, 3, 1, 1, False, "gelu", (5, 9, 31), 1),
    ("tcg3d_72_72", 3, 72, 72, 3, 1, 1, False, "gelu", (3, 6, 20), 2),
    ("tcg3d_8_24_s2", 3, 8, 24, 3, 2, 1, False, "gelu", (12, 18, 40), 1),
    ("tcg3d_odd_40_72_s2", 3, 40, 72, 3, 2, 1, False, "gelu", (3, 5, 7), 1),
    ("tcg3d_deconv_72_40", 3, 72, 40, 4, 2, 1, True, "gelu", (2, 3, 5), 1),
    ("tcg3d_deconv_40_24", 3, 40, 24, 4, 2, 1, True, "gelu", (3, 6, 20), 2),
    ("tcg3d_deconv_24_1", 3, 24, 1, 4, 2, 1, True, None, (4, 6, 13), 2),
    ("tcg2d_240_240", 2, 240, 240, 3, 1, 1, False, "gelu", (6, 20), 1),   # several channel tiles
    ("tcg2d_deconv_208_120", 2, 208, 120, 4, 2, 1, True, "gelu", (3, 10), 1),
    ("tcg2d_120_208_s2", 2, 120, 208, 3, 2, 1, False, "gelu", (6, 20), 1),
    ("tcg2d_12_12_padded_cin", 2, 12, 12, 3, 1, 1, False, "gelu", (9, 21), 1),
    ("tcg2d_s2_32_32", 2, 32, 32, 3, 2, 1, False, "gelu", (12, 40), 2),
    ("tcg2d_big", 2, 16, 24, 3, 1, 1, False, "silu", (70, 150), 1),       # many voxel tiles per CTA (ring wrap, 2 accumulators)
]


@pytest.fixture
def tcg_forced(monkeypatch):
    monkeypatch.setenv("ESM_TC_FORCE", "2")
    monkeypatch.setenv("ESM_TC", "3")
    from esmstereo_b200 import _lib
    return _lib.lib().esm_tcg_conv_launches


@pytest.mark.parametrize("case", TCG_CASES, ids=[c[0] for c in TCG_CASES])
def test_conv_streamed_tensor_core_path(case, tcg_forced):
    name, nd, cin, cout, k, stride, pad, transposed, act, sp, B = case
    ops = _ops()
    p = make_layer(cin, cout, k, nd, transposed=transposed, bn=act is not None, seed=len(name))
    x = rnd(B, cin, *sp, seed=3)
    want = ref_conv(x, p, stride, pad, transposed, act, nd)
    n0 = tcg_forced()
    got = ops.conv(x.cuda(), gpu_pack(p, stride, pad, transposed), act)
    assert tcg_forced() == n0 + 1, "layer did not take the streamed-weight tensor-core path"
    assert rel(got, want) < 2e-5, name


def test_conv_streamed_tensor_core_fusions(tcg_forced):
    ops = _ops()
    # cat(cropped 3D view, skip) -> k3 conv (three sources' worth of channel groups, 3D crop strides)
    a_full = rnd(1, 40, 4, 6, 10, seed=1)
    b = rnd(1, 40, 3, 5, 9, seed=2)
    p = make_layer(80, 40, 3, 3, seed=5)
    want = ref_conv(torch.cat((a_full[:, :, :3, :5, :9], b), 1), p, 1, 1, False, "gelu", 3)
    n0 = tcg_forced()
    got = ops.conv([a_full.cuda()[:, :, :3, :5, :9], b.cuda()], gpu_pack(p, 1, 1, False), "gelu")
    assert tcg_forced() == n0 + 1
    assert rel(got, want) < 2e-5
    # concat + residual + second activation + scale, k3
    a, b2 = rnd(2, 16, 11, 45, seed=1), rnd(2, 8, 11, 45, seed=2)
    p = make_layer(24, 16, 3, 2, bias=True, seed=3)
    res = rnd(2, 16, 11, 45, seed=4)
    want = torch.sigmoid(ref_conv(torch.cat((a, b2), 1), p, 1, 1, False, "gelu", 2) + res) * 3.0
    got = ops.conv([a.cuda(), b2.cuda()], gpu_pack(p, 1, 1, False), "gelu", residual=res.cuda(), act2="sigmoid", out_scale=3.0)
    assert tcg_forced() == n0 + 2
    assert rel(got, want) < 2e-5
    # transposed conv cropped to the skip's extent (ESMStereo.py:172), out_mul broadcast over D
    x = rnd(1, 40, 3, 6, 20, seed=7)
    p = make_layer(40, 24, 4, 3, transposed=True, seed=8)
    full = ref_conv(x, p, 2, 1, True, "gelu", 3)
    att = rnd(1, 24, 11, 39, seed=9)
    want = full[:, :, :5, :11, :39] * att.unsqueeze(2)
    got = ops.conv(x.cuda(), gpu_pack(p, 2, 1, True), "gelu", out_size=(5, 11, 39), out_mul=att.cuda())
    assert tcg_forced() == n0 + 3
    assert rel(got, want) < 2e-5


@pytest.mark.parametrize("nd,shape,bn", [(2, (1, 32, 13, 21), False), (3, (2, 24, 4, 6, 13), False), (2, (2, 16, 7, 9), True)])
def test_single_channel_deconv_subpixel_form(nd, shape, bn):
    """conv1_up (ConvTranspose k4 s2 p1 -> 1 channel) runs as a k3 conv to 2^nd phase channels + PixelShuffle."""
    from esmstereo_b200 import layers as L
    torch.manual_seed(5)
    m = L.BasicConv(shape[1], 1, deconv=True, is_3d=nd == 3, bn=bn, gelu=bn, kernel_size=(4,) * nd, padding=(1,) * nd, stride=(2,) * nd)
    with torch.no_grad():
        m.bn.running_mean.normal_(0, 0.1)
        m.bn.running_var.uniform_(0.5, 1.5)
        m.bn.weight.uniform_(0.5, 1.5)
        m.bn.bias.normal_(0, 0.1)
    m.eval()
    assert m._subpixel
    x = rnd(*shape, seed=2)
    with torch.no_grad():
        want = m.conv(x)
        if bn:
            want = F.gelu(m.bn(want))
        got = m.cuda()(x.cuda())
        assert got.shape == want.shape
        assert rel(got, want) < 2e-5
        # a fused argument other than a full-size out_size takes the generic transposed path; same answer
        crop = tuple(2 * s - 1 for s in shape[2:])
        got2 = m(x.cuda(), out_size=crop)
        assert rel(got2, want[(slice(None), slice(None)) + tuple(slice(0, c) for c in crop)]) < 2e-5


@pytest.mark.parametrize("cin,cout,shape", [(16, 64, (33, 70)), (8, 32, (12, 312)), (16, 16, (5, 9))])
def test_conv_tensor_core_pointwise_pixel_shuffle(cin, cout, shape, tc_forced):
    """UpShuffle (ESMStereo.py:265-268): 1x1 conv -> PixelShuffle(2) -> SiLU as one tcgen05 launch."""
    ops = _ops()
    p = make_layer(cin, cout, 1, 2, bn=False, bias=True, seed=11)
    x = rnd(2, cin, *shape, seed=4)
    want = F.silu(F.pixel_shuffle(ref_conv(x, p, 1, 0, False, None, 2), 2))
    n0 = tc_forced()
    got = ops.conv(x.cuda(), gpu_pack(p, 1, 0, False), None, pixel_shuffle=2, act2="silu")
    assert tc_forced() == n0 + 1, "layer did not take the tensor-core path"
    assert got.shape == want.shape
    assert rel(got, want) < 2e-5


# ---- pointwise streaming kernel (conv_pw.cu), forced on ----
@pytest.fixture
def pw_forced(monkeypatch):
    monkeypatch.setenv("ESM_TC_FORCE", "3")
    from esmstereo_b200 import _lib
    return _lib.lib().esm_pw_conv_launches


@pytest.mark.parametrize("case", TC_K1_CASES + [("pw_12_20", 2, 12, 20, (9, 21), 2), ("pw_160_32", 2, 160, 32, (6, 20), 1),
                                                  ("pw3d_80_40", 3, 80, 40, (3, 5, 9), 1)], ids=lambda c: c[0])
def test_conv_pointwise_streaming(case, pw_forced):
    name, nd, cin, cout, sp, B = case
    ops = _ops()
    p = make_layer(cin, cout, 1, nd, seed=len(name))
    x = rnd(B, cin, *sp, seed=3)
    want = ref_conv(x, p, 1, 0, False, "gelu", nd)
    n0 = pw_forced()
    got = ops.conv(x.cuda(), gpu_pack(p, 1, 0, False), "gelu")
    assert pw_forced() == n0 + 1, "layer did not take the pointwise kernel"
    assert rel(got, want) < 2e-5, name


def test_conv_pointwise_streaming_fusions(pw_forced):
    ops = _ops()
    # three concatenated sources (up_refinement.agg_0.0), residual, second activation, scale
    srcs = [rnd(2, 32, 6, 20, seed=7), rnd(2, 32, 6, 20, seed=8), rnd(2, 96, 6, 20, seed=9)]
    p = make_layer(160, 32, 1, 2, seed=6)
    res = rnd(2, 32, 6, 20, seed=10)
    want = torch.sigmoid(ref_conv(torch.cat(srcs, 1), p, 1, 0, False, "gelu", 2) + res) * 2.0
    n0 = pw_forced()
    got = ops.conv([s.cuda() for s in srcs], gpu_pack(p, 1, 0, False), "gelu", residual=res.cuda(), act2="sigmoid", out_scale=2.0)
    assert pw_forced() == n0 + 1
    assert rel(got, want) < 2e-5
    # cropped 3D view + skip (aggregation.agg_0.0), out_mul broadcast over D
    a_full, b = rnd(1, 40, 4, 6, 10, seed=1), rnd(1, 40, 3, 5, 9, seed=2)
    p = make_layer(80, 40, 1, 3, seed=5)
    att = rnd(1, 40, 5, 9, seed=3)
    want = ref_conv(torch.cat((a_full[:, :, :3, :5, :9], b), 1), p, 1, 0, False, "gelu", 3) * att.unsqueeze(2)
    got = ops.conv([a_full.cuda()[:, :, :3, :5, :9], b.cuda()], gpu_pack(p, 1, 0, False), "gelu", out_mul=att.cuda())
    assert pw_forced() == n0 + 2
    assert rel(got, want) < 2e-5
    # k1 with padding 1 (last layer of the disparity MLP, ESMStereo.py:253): the border is act(shift)
    x = rnd(1, 32, 10, 38, seed=4)
    p = make_layer(32, 32, 1, 2, seed=8)
    want = ref_conv(x, p, 1, 1, False, "gelu", 2)
    got = ops.conv(x.cuda(), gpu_pack(p, 1, 1, False), "gelu")
    assert pw_forced() == n0 + 3
    assert got.shape == want.shape and rel(got, want) < 2e-5
    # UpShuffle: 1x1 -> PixelShuffle(2) -> SiLU
    p = make_layer(16, 64, 1, 2, bn=False, bias=True, seed=11)
    x = rnd(2, 16, 33, 70, seed=4)
    want = F.silu(F.pixel_shuffle(ref_conv(x, p, 1, 0, False, None, 2), 2))
    got = ops.conv(x.cuda(), gpu_pack(p, 1, 0, False), None, pixel_shuffle=2, act2="silu")
    assert pw_forced() == n0 + 4
    assert got.shape == want.shape and rel(got, want) < 2e-5


@pytest.mark.parametrize("case", [
    # name, cin, cout, k, stride, pad, in_shape, batch
    ("thin_stem_3_32_s2", 3, 32, 3, 2, 1, (25, 63), 2),
    ("thin_dm0_1_32_k5p1", 1, 32, 5, 1, 1, (12, 40), 1),
    ("thin_ref_1_32_s2", 1, 32, 3, 2, 1, (24, 81), 1),
    ("thin_1_16_k3", 1, 16, 3, 1, 1, (9, 21), 2),
], ids=lambda c: c[0])
def test_conv_thin_input_streaming(case, pw_forced):
    """2D layers on 1 / 3 input channels (image stems, first layers on the disparity map) on the streaming kernel."""
    name, cin, cout, k, stride, pad, sp, B = case
    ops = _ops()
    p = make_layer(cin, cout, k, 2, seed=len(name))
    x = rnd(B, cin, *sp, seed=3)
    want = ref_conv(x, p, stride, pad, False, "gelu", 2)
    n0 = pw_forced()
    got = ops.conv(x.cuda(), gpu_pack(p, stride, pad, False), "gelu")
    assert pw_forced() == n0 + 1, "layer did not take the streaming kernel"
    assert got.shape == want.shape and rel(got, want) < 2e-5, name


def test_laf_sample_embed_vs_grid_sample():
    """LAFNet's scale-adaptive 3x3 sampling + embed_conv2 (k3, stride 3) + BN + ReLU, fused on the GPU, against the
    reference's own formulation: a 3h x 3w `grid_sample(align_corners=True, zeros)` image convolved with stride 3
    (ESMStereo_confidence.py:693-719; note the y offsets are +-scale and the x offsets scale * 2/(w-1))."""
    import numpy as np
    import torch.nn.functional as F
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    for (b, c, h, w) in ((1, 16, 13, 21), (2, 16, 6, 9)):
        feat = torch.randn(b, c, h, w, generator=g)
        scale = 2 * torch.sigmoid(torch.randn(b, 1, h, w, generator=g))
        wt = torch.randn(c, c, 3, 3, generator=g) * 0.1
        sc = 0.5 + torch.rand(c, generator=g)
        sh = torch.randn(c, generator=g) * 0.1
        gw, gh = np.meshgrid(np.linspace(-1, 1, w), np.linspace(-1, 1, h))
        gw = torch.tensor(gw, dtype=torch.float32).view(1, h, w, 1).expand(b, h, w, 1)
        gh = torch.tensor(gh, dtype=torch.float32).view(1, h, w, 1).expand(b, h, w, 1)
        grid = torch.cat((gw, gh), 3)
        st = scale.permute(0, 2, 3, 1)
        step_y = 2 / (w - 1)
        big = torch.zeros(b, 3 * h, 3 * w, 2)
        for iy, oy in enumerate((-1, 0, 1)):
            for ix, ox in enumerate((-1, 0, 1)):
                big[:, iy::3, ix::3, :] = grid + torch.cat((ox * step_y * st, oy * st), 3)
        samp = F.grid_sample(feat, big, mode="bilinear", padding_mode="zeros", align_corners=True)
        want = F.relu(F.conv2d(samp, wt, None, stride=3) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
        got = ops.laf_sample_embed(feat.cuda(), scale.cuda(), wt.cuda().contiguous(), sc.cuda(), sh.cuda())
        assert tuple(got.shape) == (b, c, h, w)
        assert rel(got, want) < 2e-5, (b, c, h, w)


@pytest.mark.parametrize("engine", ["resident", "streamed", "flat"])
@pytest.mark.parametrize("dist", ["all_positive", "cancelling", "peaked"])
def test_tensor_core_truncation_bias_beyond_gaussian(engine, dist, monkeypatch):
    """The tcgen05 accumulator truncates toward zero on every accumulate; the epilogues undo the EXPECTED shrink
    (1 + n * 1.25e-8, tc_common.cuh).  That constant was fitted on Gaussian data: this holds every tensor-core engine
    to the FP32-pipe gate, and its mean SIGNED error to a bias bound, on the distributions where a fixed factor could
    go wrong -- all-positive sums (post-GELU activations x positive weights: every accumulate shrinks), strongly
    cancelling sums (the accumulator is small while the addends are large) and peaked inputs."""
    import torch.nn.functional as F
    ops = _ops()
    g = torch.Generator().manual_seed(21)
    cin = cout = 40
    x = torch.randn(1, cin, 6, 12, 40, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) * 0.03
    if dist == "all_positive":
        x, w = x.abs() + 0.1, w.abs() + 0.01
    elif dist == "cancelling":
        x = 100.0 + 0.01 * x                       # large common mode ...
        w = w - w.mean(dim=(1, 2, 3, 4), keepdim=True)   # ... against zero-sum filters: outputs are O(1e-2) sums of O(100) terms
    else:
        x = 0.01 * x
        x[:, 3, 2, 5, 17] = 50.0
    want = F.conv3d(x.double(), w.double(), None, padding=1)
    if engine == "flat":
        pc = ops.pack_conv_pf(w.cuda(), [cin], 1, False, None)
        got = ops.conv_pf([ops.to_pf(x.cuda())], pc, None, out="nchw")
    else:
        monkeypatch.setenv("ESM_TC_FORCE", "1" if engine == "resident" else "2")
        pc = ops.pack_conv(w.cuda(), 1, 1, False, None, None)
        got = ops.conv(x.cuda(), pc, None)
    ref32 = ops.conv(x.cuda(), ops.pack_conv(w.cuda(), 1, 1, False, None, None), None, fp32_only=True)
    scale_ = want.abs().max().item()
    err = (got.double().cpu() - want).abs().max().item() / scale_
    err32 = (ref32.double().cpu() - want).abs().max().item() / scale_
    if dist == "cancelling":
        # the inputs themselves carry 100 * 2^-24 of representation noise per term; compare with the FP32 pipe instead of a fixed gate
        assert err <= 4 * err32 + 1e-6, (err, err32)
    else:
        assert err < 2e-5, (err, err32)
    interior = want.abs() > 0.05 * scale_
    bias = ((got.double().cpu() - want)[interior] / want[interior]).mean().item()
    assert abs(bias) < 2e-6, "mean signed relative error %g (truncation bias not compensated)" % bias


def test_regression_on_subpixel_volume_and_pixel_shuffle3d():
    """The top-2 regression reading conv1_up's 8 sub-pixel phases == the regression on the shuffled volume, bit for bit
    (indices and prediction); the shuffle kernel == torch's view/permute/reshape of the reference's ConvTranspose3d layout."""
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    for (B, D2, H2, W2) in ((1, 6, 5, 9), (2, 24, 12, 39)):
        buf = torch.randn(B, 8, D2, H2, W2 + 3, generator=g).cuda()
        y8 = buf[..., :W2]  # a pitched view, like the conv engines' outputs
        vol = ops.pixel_shuffle3d(y8)
        want = y8.contiguous().view(B, 2, 2, 2, D2, H2, W2).permute(0, 4, 1, 5, 2, 6, 3).reshape(B, 1, 2 * D2, 2 * H2, 2 * W2)
        assert torch.equal(vol, want)
        p0, i0 = ops.regression_top2(vol[:, 0].contiguous(), return_indices=True)
        p1, i1 = ops.regression_top2_subpixel(y8, return_indices=True)
        assert torch.equal(i0, i1) and torch.equal(p0, p1)


def test_disparity_publish_matches_opencv_semantics():
    """crop + medianBlur(5) (exact median, replicated border) + validity mask + round(d * 256) -> uint16
    (kitti_publisher_cuda_node.cpp:385-404), against a torch restatement."""
    import torch.nn.functional as F
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    Hp, Wp, h, w = 64, 96, 37, 61
    d = torch.rand(Hp, Wp, generator=g) * 230.0 - 10.0  # some values outside (0, 192)
    crop = d[:h, :w]
    pad = F.pad(crop[None, None], (2, 2, 2, 2), mode="replicate")
    med = pad.unfold(2, 5, 1).unfold(3, 5, 1).reshape(h, w, 25).sort(-1).values[..., 12]
    med = torch.where((med > 0) & (med < 192.0), med, torch.zeros_like(med))
    want = torch.clamp(torch.round(med * 256.0), 0, 65535).to(torch.int32)
    got = ops.disparity_publish_u16(d.cuda(), (h, w), 192.0, 256.0).cpu().to(torch.int32)
    assert torch.equal(got, want)


@pytest.mark.parametrize("cout,act,shape,batch", [(32, "relu6", (96, 160), 2), (32, "gelu", (75, 131), 1), (16, "gelu", (130, 66), 1)])
def test_image_stem_kernel(cout, act, shape, batch):
    """conv_stem / stem_2[0] (3 -> C, k3 s2 p1 + BN + activation; ESMStereo.py:49,529-533) on the dedicated kernel
    (conv_stem3.cu): even / odd sizes, partial tiles, both channel widths, a strided (pitched) input view."""
    ops = _ops()
    p = make_layer(3, cout, 3, 2, seed=cout + shape[0])
    buf = rnd(batch, 3, shape[0], shape[1] + 5, seed=2)
    x = buf[..., : shape[1]]
    want = ref_conv(x.contiguous(), p, 2, 1, False, act, 2)
    got = ops.conv(buf.cuda()[..., : shape[1]], gpu_pack(p, 2, 1, False), act)
    assert rel(got, want) < 2e-6, (cout, shape)
