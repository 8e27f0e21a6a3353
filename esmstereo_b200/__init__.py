"""esmstereo_b200 -- B200-native (sm_100a) implementation of ESMStereo's feature-to-disparity hot path.

Drop-in for the reference's `models` package on that path:

    from esmstereo_b200 import __models__
    model = __models__["ESMStereo"](192, True, False, "efficientnet_b2", 4).cuda().eval()
    disp = model(left, right, train_status=False)[-1]

The hot path has no CPU / PyTorch fallback: it needs esmstereo_b200/csrc/libesm_b200.so
(`python -m esmstereo_b200.build`) and a CUDA device.
"""
from .model import (ESMStereo, ESMStereo_confidence, ESMStereo_trt, GraphedStereo, StereoPipeline,  # noqa: F401
                    __models__, load_reference_checkpoint)
from .ops import (build_gwc_volume, build_norm_correlation_volume, disparity_regression,  # noqa: F401
                  regression_topk)

__all__ = ["ESMStereo", "ESMStereo_trt", "ESMStereo_confidence", "GraphedStereo", "StereoPipeline", "__models__", "load_reference_checkpoint",
           "build_gwc_volume", "build_norm_correlation_volume", "disparity_regression", "regression_topk"]
