"""ESMStereo model family behind the reference's own Python API (drop-in boundary):

    __models__[name](maxdisp, gwc, norm_correlation, backbone, cv_scale[, device])
    ESMStereo(left, right, train_status) -> [disp [B,H,W]]          (models/ESMStereo.py:512,638)
    ESMStereo_trt(left, right)           -> disp [B,H,W]            (models/ESMStereo_trt.py:638,735)
    ESMStereo_confidence(left, right)    -> (disp, conf) [B,H,W]    (models/ESMStereo_confidence.py:876,974)

Parameter / buffer names are the reference's, so checkpoints load by key (test_kitti.py:57-61) and the
modules survive nn.DataParallel(...).cuda().eval().  The 2D feature side is PyTorch/cuDNN
(`feature2d`); from the matching descriptors on -- cost volume, 3D hourglass, regression, ShuffleMixer
upsampling, confidence head -- every op is a libesm_b200 kernel (`layers`, `ops`).
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import layers as L
from . import ops
from .feature2d import ConvThenPlain, FeatUp, Feature, TorchBasicConv, fused, image_stem

_STEM_CH = {4: [(3, 32), (32, 48)], 8: [(3, 32), (32, 48), (48, 64)], 16: [(3, 16), (16, 24), (24, 32), (32, 40)]}
_DESC_IN = {4: 96, 8: 160, 16: 136}
_ADD_CH = {4: 16, 8: 8, 16: 4}


@contextlib.contextmanager
def _exact_fp32(enabled: bool):
    """fp32 mode: keep cuDNN / cuBLAS out of TF32 on the 2D side so the descriptors feeding the
    cost volume match the CPU reference (hard part 2)."""
    if not enabled:
        yield
        return
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True  # cuDNN's transposed convs otherwise pick atomics-based kernels
    try:
        yield
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
         torch.backends.cudnn.deterministic) = old


class _ESMStereoBase(nn.Module):
    def __init__(self, maxdisp: int, gwc: bool = False, norm_correlation: bool = True, backbone: str = "efficientnet_b2",
                 cv_scale: int = 4, confidence: bool = False) -> None:
        super().__init__()
        self.maxdisp, self.vol_size = maxdisp, cv_scale
        self.gwc, self.norm_correlation, self.backbone = gwc, norm_correlation, backbone
        self.exact_fp32 = True      # engine "torch" only: keep cuDNN out of TF32 on the 2D feature side (parity-grade)
        self.fuse_volume = True     # gwc volume generated inside group_stem (never written to HBM)
        self.fork_streams = os.environ.get("ESM_FORK", "1") != "0"
        self._side_streams = {}
        self.capture: Optional[Dict[str, torch.Tensor]] = None  # set to {} to record hot-path stages (tests)
        if cv_scale not in (4, 8, 16):
            # the reference hits a misspelt `pirnt(...)` here -> NameError (ESMStereo.py:599)
            raise NameError("Choose the cost volume resolution: 4, 8, 16")
        self.feature = Feature(backbone)
        # 2D feature side: "esm" = the fused direct-conv kernels (stand-in backbone), "torch" = cuDNN modules
        self.feature_engine = "esm" if self.feature.esm_capable else "torch"
        if confidence and self.feature_engine == "esm":
            self.feature_engine = "esm_fp32"  # everything upstream of the confidence head's cost tower stays on the FP32 pipe (below)
        if cv_scale in (4, 8):
            self.feature_up = FeatUp(self.feature.chans, cv_scale)
        for (cin, cout), n in zip(_STEM_CH[cv_scale], (2, 4, 8, 16)):
            setattr(self, "stem_%d" % n, image_stem(cin, cout))
        self.conv = TorchBasicConv(_DESC_IN[cv_scale], 64, kernel_size=3, padding=1, stride=1)
        self.desc = nn.Conv2d(64, 64, kernel_size=1, padding=0, stride=1)
        self._desc_pc = L._Packed()
        if cv_scale == 16:
            self.conv_f2 = TorchBasicConv(96, 32, kernel_size=3, padding=1, stride=1)
            self.conv_f0 = TorchBasicConv(16, 24, kernel_size=3, padding=1, stride=1)
        k3 = dict(deconv=False, is_3d=True, bn=True, gelu=True, kernel_size=3, padding=1, stride=1)
        if norm_correlation:
            print("Cost volumes: norm correlation")
            if cv_scale == 16:
                self.semantic = ConvThenPlain(96, 32, 8)
            self.corr_stem = L.BasicConv(1, 8, **k3)
        if gwc:
            print("Cost volumes: gwc ")
            if cv_scale == 16:
                self.semantic = ConvThenPlain(96, 64, 32)
            self.num_groups = 32
            self.group_stem = L.BasicConv(self.num_groups, 8, **k3)
        self.agg = L.BasicConv(8, 8, **k3)
        self.upsample_module = {4: L.upsample4, 8: L.upsample8, 16: L.upsample16}[cv_scale]()
        if confidence and cv_scale == 16:
            self.confidence_net = L.LAFNet_ESM(16)
        self.aggregation_out = L.aggregation(8, _ADD_CH[cv_scale])
        if confidence:
            # LAFNet's cost tower is softmax(-100 * cost / |cost|) (ESMStereo_confidence.py:575): it amplifies the cost
            # volume's rounding error a hundredfold.  The split-TF32 tensor-core engines are ~3x less accurate than the
            # FP32 pipe (the tensor core truncates its fp32 accumulator), so this model keeps the 3D path on the FP32
            # pipe; at cv_scale 16 that path is 1.5 GFLOP, so nothing is lost.
            for mod in [self.agg, self.aggregation_out] + [getattr(self, n) for n in ("corr_stem", "group_stem") if hasattr(self, n)]:
                for sub in mod.modules():
                    if isinstance(sub, L.BasicConv):
                        sub.fp32_only = True

    # ------------------------------------------------------------------ 2D side (PyTorch)
    def _features_2d(self, left: torch.Tensor, right: torch.Tensor):
        """Returns (features_left list, stems of the left image, match_left, match_right, att).
        Both images go through the shared layers as one batch (ESMStereo.py:640-697)."""
        B = left.shape[0]
        both = ops.cat_batch(left, right)
        eng = self.feature_engine
        with _exact_fp32(self.exact_fp32 and eng == "torch"):
            # the image stems depend on the images only: they run on a side stream next to the backbone + FeatUp chain
            # (fork / join by events, captured as parallel branches of the CUDA graph); ESM_FORK=0 keeps one stream
            fork = self.fork_streams and both.is_cuda
            cur = torch.cuda.current_stream(both.device) if fork else None
            if fork:
                side = self._side_streams.get(both.device)
                if side is None:
                    side = self._side_streams[both.device] = torch.cuda.Stream(device=both.device)
                side.wait_stream(cur)
            with (torch.cuda.stream(side) if fork else contextlib.nullcontext()):
                stems = [self.stem_2(both, eng)]
                for n in (4, 8, 16):
                    if hasattr(self, "stem_%d" % n):
                        stems.append(getattr(self, "stem_%d" % n)(stems[-1], eng))
            feats = self.feature(both, eng if self.feature.esm_capable else "torch")
            if self.vol_size in (4, 8):
                feats = self.feature_up(feats, eng)
            if fork:
                cur.wait_stream(side)
                for t in stems:
                    t.record_stream(cur)
            coarse = {4: feats[0], 8: feats[1], 16: feats[3]}[self.vol_size]
            match = fused(self._desc_pc, self.desc, None, self.conv([coarse, stems[-1]], eng), None, eng)
            fl = [f[:B] for f in feats]
            att = self.semantic(fl[3], eng) if self.vol_size == 16 else None
            extra = None
            if self.vol_size == 16:
                extra = (self.conv_f2(fl[3], eng), self.conv_f0(fl[0], eng))
        return fl, [s[:B] for s in stems], match[:B].contiguous(), match[B:].contiguous(), att, extra

    # ------------------------------------------------------------------ hot path (libesm_b200)
    def _aggregate_cost(self, mL, mR, att, want_volume: bool = False) -> torch.Tensor:
        """Descriptors -> aggregated cost [B,D,h,w]: volume, stem, agg, 3D hourglass (ESMStereo.py:700-716)."""
        D = self.maxdisp // self.vol_size
        if self.norm_correlation:
            vol = ops.build_norm_correlation_volume(mL, mR, D)
            vol = self.corr_stem(vol, out_mul=att)  # `corr_stem(volume) * att` for cv16 (:703)
        if self.gwc:
            if self.fuse_volume:
                vol = self.group_stem([mL, mR], gwc_disp=D, in_mul=att)  # group_stem(volume * att) (:711)
            else:
                vol = ops.build_gwc_volume(mL, mR, D, self.num_groups)
                vol = self.group_stem(vol, in_mul=att)
        stem = vol
        vol = self.agg(vol)
        if self.vol_size == 4 and self.capture is None and not want_volume:
            # cv4: the top-2 regression reads conv1_up's sub-pixel phases directly -- the shuffled volume is never written
            y8 = self.aggregation_out(vol, keep_subpixel=True)
            if y8.shape[1] == 8:
                return y8
            return y8[:, 0]
        cost = self.aggregation_out(vol)[:, 0]
        if self.capture is not None:
            self.capture.update(match_left=mL, match_right=mR, stem=stem, agg=vol, cost=cost)
        return cost

    def _disparity_from_cost(self, cost, fl, stems, extra, want_scales: bool):
        """Aggregated cost -> (initial disparity, upsampled scales): regression + ShuffleMixer
        upsampling (ESMStereo.py:718-745)."""
        s, D = self.vol_size, self.maxdisp // self.vol_size
        final = 1.0 if want_scales else 4.0  # every output is *4 regardless of scale (:737-745)
        if s == 4:
            init = ops.regression_top2_subpixel(cost) if cost.dim() == 5 else ops.regression_top2(cost)
            scales = self.upsample_module(fl[1], fl[0], stems[0], init, out_scale=final)
        elif s == 8:
            init = ops.disparity_regression(cost, D).unsqueeze(1)
            scales = self.upsample_module(fl[2], fl[1], fl[0], stems[0], init, out_scale=final)
        else:
            init = ops.disparity_regression(cost, D).unsqueeze(1)
            f2, f0 = extra
            scales = self.upsample_module(fl[2], f2, fl[1], f0, init, out_scale=final)
        if self.capture is not None:
            self.capture.update(init_pred=init)
            if s == 4:
                self.capture["top2_idx"] = ops.regression_top2(cost, return_indices=True)[1]
        return init, scales

    def _run(self, left: torch.Tensor, right: torch.Tensor, want_scales: bool, want_conf: bool):
        if not (left.is_cuda and right.is_cuda):
            raise RuntimeError("esmstereo_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.training:
            raise RuntimeError("esmstereo_b200 implements the inference path only: call model.eval() first")
        with torch.no_grad():
            left, right = left.float().contiguous(), right.float().contiguous()
            fl, stems, mL, mR, att, extra = self._features_2d(left, right)
            cost = self._aggregate_cost(mL, mR, att, want_volume=want_conf)
            init, scales = self._disparity_from_cost(cost, fl, stems, extra, want_scales)
            conf = None
            if want_conf:
                conf = self.confidence_net(cost, init, mL, fl[3], fl[1]).squeeze(1)
            if want_scales:
                return [t.squeeze(1) * 4 for t in scales], conf
            return [scales[0].squeeze(1)], conf


class ESMStereo(_ESMStereoBase):
    def __init__(self, maxdisp: int, gwc: bool = False, norm_correlation: bool = True, backbone: str = "efficientnet_b2",
                 cv_scale: int = 4) -> None:
        super().__init__(maxdisp, gwc, norm_correlation, backbone, cv_scale)

    def forward(self, left: torch.Tensor, right: torch.Tensor, train_status: bool) -> List[torch.Tensor]:
        return self._run(left, right, bool(train_status), False)[0]


class ESMStereo_trt(_ESMStereoBase):
    def __init__(self, maxdisp: int, gwc: bool = False, norm_correlation: bool = True, backbone: str = "efficientnet_b2",
                 cv_scale: int = 4) -> None:
        super().__init__(maxdisp, gwc, norm_correlation, backbone, cv_scale)

    def forward(self, left: torch.Tensor, right: torch.Tensor) -> torch.Tensor:
        return self._run(left, right, False, False)[0][0]


class ESMStereo_confidence(_ESMStereoBase):
    def __init__(self, maxdisp: int, gwc: bool = False, norm_correlation: bool = True, backbone: str = "efficientnet_b2",
                 cv_scale: int = 4, device=torch.device("cuda")) -> None:
        super().__init__(maxdisp, gwc, norm_correlation, backbone, cv_scale, confidence=True)
        self.device = device

    def forward(self, left: torch.Tensor, right: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.vol_size != 16:
            # the reference only assigns conf_out under `if self.vol_size == 16` (:966-974)
            raise UnboundLocalError("conf_out is only produced for cv_scale == 16 (ESMStereo_confidence.py:966-974)")
        disp, conf = self._run(left, right, False, True)
        return disp[0], conf


def load_reference_checkpoint(model: nn.Module, state_dict: Dict[str, torch.Tensor], require_backbone: bool = True) -> List[str]:
    """The reference's checkpoint idiom (test_kitti.py:57-61): keep the keys the model has, update, load -- with or
    without the `module.` prefix nn.DataParallel adds -- but, unlike it, NOT silently: every model tensor the
    checkpoint does not provide is returned, and a backbone (`feature.*`) left unfilled raises (that is what a
    stand-in backbone, or a mismatched timm version, looks like)."""
    model_dict = model.state_dict()
    prefixed = any(k.startswith("module.") for k in model_dict)
    fixed = {}
    for k, v in state_dict.items():
        bare = k[len("module."):] if k.startswith("module.") else k
        fixed[("module." + bare) if prefixed else bare] = v
    pre = {k: v for k, v in fixed.items() if k in model_dict}
    missing = [k for k in model_dict if k not in pre and not k.endswith("num_batches_tracked")]
    bad = [k for k in missing if ".feature." in "." + k]
    if require_backbone and bad:
        raise RuntimeError("checkpoint leaves %d backbone tensors unfilled (e.g. %s): the model's `feature.*` names do not match the "
                           "checkpoint's -- stand-in backbone, or a different timm naming scheme" % (len(bad), bad[0]))
    model_dict.update(pre)
    model.load_state_dict(model_dict)
    return missing


__models__ = {
    "ESMStereo": ESMStereo,
    "ESMStereo_trt": ESMStereo_trt,
    "ESMStereo_confidence": ESMStereo_confidence,
}


class GraphedStereo:
    """CUDA-graph replay of one model at one input shape (the batch-1 latency path is launch-bound:
    ~150 kernels of 5-100 us each).  Inputs are copied into static buffers, outputs are views of
    static buffers that the next `__call__` overwrites.

        g = GraphedStereo(model, (1, 3, 384, 1248));  disp = g(left, right)
    """

    def __init__(self, model: nn.Module, shape, warmup: int = 3, **fwd_kwargs) -> None:
        self.model, self.kw = model, fwd_kwargs
        dev = next(model.parameters()).device
        self.left = torch.zeros(shape, device=dev)
        self.right = torch.zeros(shape, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # packs weights, sets func attributes, warms cuDNN heuristics
                self.model(self.left, self.right, **self.kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self.model(self.left, self.right, **self.kw)

    def __call__(self, left: torch.Tensor, right: torch.Tensor):
        self.left.copy_(left, non_blocking=True)
        self.right.copy_(right, non_blocking=True)
        self.graph.replay()
        return self.out


class StereoPipeline:
    """Host-to-host serving loop around a `GraphedStereo`: pinned host images in, pinned host disparity out.  The H2D
    copy of pair i+1 and the D2H copy of pair i-1 run on two copy streams and overlap the graph replay of pair i:
    the replay's output is first moved to a per-slot device buffer on the compute stream (so the next replay can start
    at once), the D2H copy of that buffer then goes on the copy stream.

        pipe = StereoPipeline(GraphedStereo(model, shape, train_status=False))
        pipe.submit(left_pinned, right_pinned); ...; disp_host = pipe.result()   # results come back in order

    `result()` returns the slot's pinned buffer, which is overwritten `depth` submits later: copy it if you keep it."""

    def __init__(self, graphed: GraphedStereo, depth: int = 2, pick=lambda out: out[-1]) -> None:
        self.g, self.depth, self.pick = graphed, depth, pick
        dev = graphed.left.device
        self.copy_stream = torch.cuda.Stream(device=dev)   # host -> device
        self.down_stream = torch.cuda.Stream(device=dev)   # device -> host: its own stream, or the H2D of pair i+1 would queue behind
        example = pick(graphed.out)                        # the D2H of pair i, which waits for replay i

        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                left=torch.empty_like(graphed.left), right=torch.empty_like(graphed.right),
                dev_out=torch.empty_like(example),
                out=torch.empty(example.shape, dtype=example.dtype).pin_memory(),
                staged=torch.cuda.Event(), consumed=torch.cuda.Event(), replayed=torch.cuda.Event(), done=torch.cuda.Event()))
        self.submitted = self.returned = 0

    def submit(self, left_host: torch.Tensor, right_host: torch.Tensor) -> None:
        if self.submitted - self.returned >= self.depth:
            raise RuntimeError("StereoPipeline: %d results outstanding, call result() first" % self.depth)
        slot = self.slots[self.submitted % self.depth]
        cur = torch.cuda.current_stream(self.g.left.device)
        with torch.cuda.stream(self.copy_stream):
            if self.submitted >= self.depth:
                self.copy_stream.wait_event(slot["consumed"])  # staging buffers were read by the replay `depth` steps ago
            slot["left"].copy_(left_host, non_blocking=True)
            slot["right"].copy_(right_host, non_blocking=True)
            slot["staged"].record(self.copy_stream)
        cur.wait_event(slot["staged"])
        if self.submitted >= self.depth:
            cur.wait_event(slot["done"])  # the D2H copy of this slot's previous result has left dev_out
        out = self.pick(self.g(slot["left"], slot["right"]))
        slot["consumed"].record(cur)
        slot["dev_out"].copy_(out, non_blocking=True)
        slot["replayed"].record(cur)
        with torch.cuda.stream(self.down_stream):
            self.down_stream.wait_event(slot["replayed"])
            slot["out"].copy_(slot["dev_out"], non_blocking=True)
            slot["done"].record(self.down_stream)
        self.submitted += 1

    def result(self) -> torch.Tensor:
        if self.returned >= self.submitted:
            raise RuntimeError("StereoPipeline: nothing submitted")
        slot = self.slots[self.returned % self.depth]
        slot["done"].synchronize()
        self.returned += 1
        return slot["out"]
