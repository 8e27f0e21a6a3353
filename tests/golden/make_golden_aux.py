"""Golden vectors for the auxiliary operators, produced by the REAL reference code
(read-only at /root/reference) and by torchvision / PIL exactly as the reference's scripts call them.  Run in the
build container only:

    python tests/golden/make_golden_aux.py        # writes tests/golden/aux_ops.npz

Inputs are regenerated from seeds by `aux_inputs()` below (imported by the tests), so only outputs are stored.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def aux_inputs():
    g = torch.Generator().manual_seed(1234)
    L = torch.randn(2, 12, 5, 17, generator=g)
    R = torch.randn(2, 12, 5, 17, generator=g)
    img = (torch.rand(37, 61, 3, generator=g) * 255).to(torch.uint8)          # h=37, w=61 -> padded to 64 x 64
    disp = torch.rand(1, 64, 64, generator=g) * 190.0
    disp[0, 40, 10] = 3.001953125   # * 256 = 768.5 -> ties to even (768)
    disp[0, 41, 10] = 3.005859375   # * 256 = 769.5 -> 770
    return L, R, img, disp


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    sys.path.insert(0, os.path.join(HERE, "timm_shim"))
    sys.path.insert(1, "/root/reference")
    sub = importlib.import_module("models.submodule")
    from PIL import Image
    from torchvision import transforms
    L, R, img, disp = aux_inputs()
    out = {"concat": sub.build_concat_volume(L, R, 6).numpy(), "substract": sub.build_substract_volume(L, R, 6, 4).numpy(),
           "gwc_norm": sub.build_gwc_volume_norm(L, R, 6, 4).numpy()}
    # test_kitti.py:93-106
    pil = Image.fromarray(img.numpy())
    w, h = pil.size
    m = 32
    wi, hi = (w // m + 1) * m, (h // m + 1) * m
    tr = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    out["pre_test_kitti"] = tr(pil.crop((w - wi, h - hi, w, h))).unsqueeze(0).numpy()
    out["pre_size"] = np.array([hi, wi])
    # datasets/kitti_dataset.py:145-160 (with 64 x 64 in place of 384 x 1248)
    t = tr(pil).numpy()
    out["pre_kitti_dataset"] = np.pad(t, ((0, 0), (hi - h, 0), (0, wi - w)), mode="constant", constant_values=0)[None]  # np.lib.pad in numpy < 2
    # test_kitti.py:114,127 and save_disp.py:83-88
    d = disp.numpy()
    out["post_test_kitti"] = np.round(d[:, hi - h:, wi - w:] * 256).astype(np.uint16)
    out["post_kitti_dataset"] = np.round(d[:, hi - h:, :-(wi - w)] * 256).astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "aux_ops.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
