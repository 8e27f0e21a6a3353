"""Auxiliary operators either side of the hot path: image pre-processing, disparity
post-processing, the reference's unused volume builders.  Golden vectors come from the real reference code /
torchvision / PIL (tests/golden/make_golden_aux.py -> aux_ops.npz); the oracle restatements are held to them on CPU,
the CUDA kernels on the GPU -- bit-exact except the squared-difference volume (summation order, 1e-6)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_aux import aux_inputs  # noqa: E402

from oracle.esm_oracle import EsmOracle, disparity_to_uint16, preprocess_images  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_ops.npz"))


def test_oracle_matches_reference_aux_goldens():
    L, R, img, disp = aux_inputs()
    o = EsmOracle({}, 192)
    assert np.array_equal(o.concat_volume(L, R, 6).numpy(), GOLD["concat"])
    assert np.abs(o.substract_volume(L, R, 6, 4).numpy() - GOLD["substract"]).max() < 1e-6
    assert np.abs(o.gwc_volume_norm(L, R, 6, 4).numpy() - GOLD["gwc_norm"]).max() < 1e-6
    size = tuple(int(v) for v in GOLD["pre_size"])
    assert np.array_equal(preprocess_images(img[None], size, "test_kitti").numpy(), GOLD["pre_test_kitti"])
    assert np.array_equal(preprocess_images(img[None], size, "kitti_dataset").numpy(), GOLD["pre_kitti_dataset"])
    assert np.array_equal(disparity_to_uint16(disp, (37, 61), "test_kitti").numpy(), GOLD["post_test_kitti"].astype(np.int32))
    assert np.array_equal(disparity_to_uint16(disp, (37, 61), "kitti_dataset").numpy(), GOLD["post_kitti_dataset"].astype(np.int32))


@pytest.mark.gpu
def test_gpu_volume_variants_match_reference():
    from esmstereo_b200 import ops
    L, R, _, _ = aux_inputs()
    got = ops.build_concat_volume(L.cuda(), R.cuda(), 6).cpu().numpy()
    assert np.array_equal(got, GOLD["concat"])
    got = ops.build_substract_volume(L.cuda(), R.cuda(), 6, 4).cpu().numpy()
    assert np.abs(got - GOLD["substract"]).max() < 1e-6
    got = ops.build_gwc_volume_norm(L.cuda(), R.cuda(), 6, 4).cpu().numpy()
    assert np.abs(got - GOLD["gwc_norm"]).max() < 1e-6
    # odd width (scalar tail of the 4-wide rows), maxdisp > W
    o = EsmOracle({}, 192)
    g = torch.Generator().manual_seed(3)
    L2, R2 = torch.randn(1, 6, 3, 7, generator=g), torch.randn(1, 6, 3, 7, generator=g)
    assert torch.equal(ops.build_concat_volume(L2.cuda(), R2.cuda(), 9).cpu(), o.concat_volume(L2, R2, 9))
    assert (ops.build_substract_volume(L2.cuda(), R2.cuda(), 9, 3).cpu() - o.substract_volume(L2, R2, 9, 3)).abs().max() < 1e-6
    assert (ops.build_gwc_volume_norm(L2.cuda(), R2.cuda(), 9, 3).cpu() - o.gwc_volume_norm(L2, R2, 9, 3)).abs().max() < 1e-6
    # two channels per group (the models' gwc geometry): every operation is then the reference's, bit for bit up to the mean
    L3, R3 = torch.randn(1, 64, 6, 40, generator=g), torch.randn(1, 64, 6, 40, generator=g)
    assert (ops.build_gwc_volume_norm(L3.cuda(), R3.cuda(), 12, 32).cpu() - o.gwc_volume_norm(L3, R3, 12, 32)).abs().max() < 1e-6
    with pytest.raises(AssertionError):
        ops.build_substract_volume(L2.cuda(), R2.cuda(), 4, 4)  # C % G != 0, submodule.py:107
    with pytest.raises(AssertionError):
        ops.build_gwc_volume_norm(L2.cuda(), R2.cuda(), 4, 4)  # submodule.py:165


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["test_kitti", "kitti_dataset"])
def test_gpu_pre_post_processing_bit_exact(mode):
    from esmstereo_b200 import ops
    _, _, img, disp = aux_inputs()
    size = tuple(int(v) for v in GOLD["pre_size"])
    got = ops.preprocess_images(img[None].cuda(), size, mode).cpu().numpy()
    assert np.array_equal(got, GOLD["pre_" + mode])
    u16 = ops.disparity_to_uint16(disp.cuda(), (37, 61), mode).cpu()
    assert np.array_equal(u16.numpy().astype(np.int32), GOLD["post_" + mode].astype(np.int32))
    # KITTI size, batch 2, against the oracle
    g = torch.Generator().manual_seed(9)
    imgs = (torch.rand(2, 375, 1242, 3, generator=g) * 255).to(torch.uint8)
    want = preprocess_images(imgs, (384, 1248), mode)
    assert torch.equal(ops.preprocess_images(imgs.cuda(), (384, 1248), mode).cpu(), want)
    d = torch.rand(2, 384, 1248, generator=g) * 191.0
    want16 = disparity_to_uint16(d, (375, 1242), mode)
    assert torch.equal(ops.disparity_to_uint16(d.cuda(), (375, 1242), mode).cpu().to(torch.int32), want16)
    with pytest.raises(TypeError):
        ops.preprocess_images(imgs, (384, 1248), mode)  # CPU tensor: no fallback
