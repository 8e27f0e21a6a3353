"""Name-keyed deterministic weights.

The reference ships no checkpoints (`/root/reference/.gitignore:1-2`), so parity runs use random
weights.  Module construction order differs between the reference and this package, therefore the
values are derived from each tensor's NAME (crc32) + a seed, never from global RNG state: any two
state_dicts with the same keys and shapes get bit-identical contents.
"""
import math
import zlib

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def fill_deterministic(state_dict, seed: int = 0):
    """In-place fill of every tensor in `state_dict`; returns it.  Scales are chosen so that
    activations neither vanish nor explode before BN calibration (see oracle `calibrate`)."""
    for key in sorted(state_dict.keys()):
        t = state_dict[key]
        g = _gen(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t.zero_()
        elif leaf == "running_mean":
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        elif leaf == "running_var":
            t.copy_(0.5 + torch.rand(t.shape, generator=g))
        elif t.dim() >= 3:  # conv / deconv kernels
            fan_in = t[0].numel() if t.shape[0] > 0 else 1
            # transposed convs store [Cin, Cout, k...]; their effective fan-in per output is
            # Cin * (k/stride)^n, close enough to t[:,0].numel()/2^n; a common gain keeps it simple
            std = math.sqrt(2.0 / max(fan_in, 1))
            t.copy_(std * torch.randn(t.shape, generator=g))
        elif leaf == "weight":  # BN gamma / LayerNorm gain
            t.copy_(0.5 + torch.rand(t.shape, generator=g))
        elif leaf == "bias":
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        else:
            t.copy_(torch.randn(t.shape, generator=g))
    return state_dict


def synthetic_pair(batch: int, height: int, width: int, shift: int = 7, seed: int = 0):
    """SURVEY.md section 8(d) synthetic inputs: right = left rolled by `shift` px + 5% noise, so a
    real correspondence exists.  Returns CPU fp32 tensors [B,3,H,W]."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    left = torch.randn(batch, 3, height, width, generator=g)
    # smooth a little so features are not pure white noise
    left = torch.nn.functional.avg_pool2d(left, 3, 1, 1) * 2.0
    right = torch.roll(left, -shift, dims=3) + 0.05 * torch.randn(batch, 3, height, width, generator=g)
    return left.contiguous(), right.contiguous()
