"""Benchmark of the ESMStereo hot path (BASELINE.json: stereo pairs/sec @384x1248 maxdisp192; cost-volume HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config A|B|C|E] [--batch b] [--breakdown]

One step = one forward of the configuration's model over one synthetic batch per GPU (default config B =
BASELINE.json configs[1]: ESMStereo-L cv4 gwc, one KITTI-shaped 384x1248 pair); pairs shard by rank with no data-path
collective (weak scaling), and the per-rank disparities are all-gathered over NCCL on a side stream as the
reference-side "gather of outputs".  Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port
instead (the Python reference cannot travel to the GPU box).

Configurations (BASELINE.json `configs`): A = 256x512 pair (the reference's CPU-runnable case), B = 384x1248 pair
(latency path, the headline), C = 8 x 544x960 pairs (throughput path), E = ESMStereo_confidence 992x1472 cv16;
D (the batch sweep across GPUs) is `--config B --batch b --gpus N`.
"""
import argparse
import contextlib
import ctypes
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the synthetic workload (and the CPU oracle it is compared with) is defined on the structural stand-in backbone, as the
# workload strings say; the package default is the timm-compatible definition (esmstereo_b200/backbone.py)
os.environ.setdefault("ESM_BACKBONE", "standin")
import warnings  # noqa: E402
warnings.filterwarnings("ignore", message="esmstereo_b200: ESM_BACKBONE=standin")

MAXDISP = 192
CONFIGS = {
    "A": dict(model="ESMStereo", backbone="efficientnet_b2", cv=4, H=256, W=512, batch=1,
              metric="stereo pairs/sec @256x512 maxdisp192",
              workload="ESMStereo-L (cv_scale 4, gwc, efficientnet_b2 stand-in backbone) 256x512 pair, batch 1 per GPU, maxdisp 192, fp32"),
    "B": dict(model="ESMStereo", backbone="efficientnet_b2", cv=4, H=384, W=1248, batch=1,
              metric="stereo pairs/sec @384x1248 maxdisp192",
              workload="ESMStereo-L (cv_scale 4, gwc, efficientnet_b2 stand-in backbone) 384x1248 pair, batch 1 per GPU, maxdisp 192, fp32"),
    "C": dict(model="ESMStereo", backbone="efficientnet_b2", cv=4, H=544, W=960, batch=8,
              metric="stereo pairs/sec @544x960 maxdisp192 batch 8",
              workload="ESMStereo-L (cv_scale 4, gwc, efficientnet_b2 stand-in backbone) 544x960 pairs, batch 8 per GPU, maxdisp 192, fp32"),
    "E": dict(model="ESMStereo_confidence", backbone="mobilenetv2_100", cv=16, H=992, W=1472, batch=1,
              metric="stereo pairs/sec @992x1472 maxdisp192 with confidence",
              workload="ESMStereo_confidence (cv_scale 16, gwc, mobilenetv2_100 stand-in backbone) 992x1472 pair, batch 1 per GPU, maxdisp 192, fp32"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="pairs per step per GPU (default: the configuration's)")
    ap.add_argument("--breakdown", action="store_true", help="also print a per-operator timing table to stderr")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity check, the GPU-eager baseline and the kernel rooflines")
    a = ap.parse_args()
    a.cfg = dict(CONFIGS[a.config])
    if a.batch > 0:
        a.cfg["batch"] = a.batch
        a.cfg["workload"] = a.cfg["workload"].replace("batch %d per GPU" % CONFIGS[a.config]["batch"], "batch %d per GPU" % a.batch)
    return a


def build_weights(cfg, seed=0):
    import torch  # noqa: F401
    from esmstereo_b200 import __models__
    from esmstereo_b200.weights import fill_deterministic
    with contextlib.redirect_stdout(io.StringIO()):
        model = __models__[cfg["model"]](MAXDISP, True, False, cfg["backbone"], cfg["cv"])
    sd = fill_deterministic(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    return model, sd


def make_oracle(cfg, sd, device="cpu"):
    from oracle.esm_oracle import EsmOracle
    return EsmOracle(sd, MAXDISP, True, False, cfg["backbone"], cfg["cv"], confidence=cfg["model"] == "ESMStereo_confidence", device=device)


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference forward, all host threads
# ----------------------------------------------------------------------------------------------
def cpu_forward_rate(cfg, sd, budget_s, min_iters=2):
    """Bounded sample of the same workload on the host cores: forwards of ONE pair of the configuration's shape."""
    import torch
    from esmstereo_b200.weights import synthetic_pair
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = make_oracle(cfg, sd)
    left, right = synthetic_pair(1, cfg["H"], cfg["W"], shift=23, seed=0)
    orc(left, right)  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_iters or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        orc(left, right)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return 1.0 / med, cores, len(times), med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    _, sd = build_weights(cfg)
    import torch
    from esmstereo_b200.weights import synthetic_pair
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = make_oracle(cfg, sd)
    B = cfg["batch"]
    left, right = synthetic_pair(B, cfg["H"], cfg["W"], shift=23, seed=0)
    t0 = time.perf_counter()
    orc(left, right)  # first warm-up, also the estimate that bounds the sample
    est = time.perf_counter() - t0
    # honour --warmup / --steps unless the whole run would exceed ~4 minutes of CPU work
    budget = 240.0
    warm = max(1, args.warmup)
    steps = max(1, args.steps)
    if (warm + steps) * est > budget:
        warm = max(1, min(warm, int(0.2 * budget / est)))
        steps = max(1, int((budget - warm * est) / est))
    for _ in range(warm - 1):
        orc(left, right)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc(left, right)
    dt = time.perf_counter() - t0
    value = steps * B / dt
    sample = "%d forwards of one batch of %d %dx%d pair(s) (%d steps asked), torch-CPU fp32, %d threads" % (
        steps, B, cfg["H"], cfg["W"], args.steps, cores)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["workload"]},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampling
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.marks = index, [], None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("ESM_CLOCK_MS", "100")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def wait_first(self, timeout_s):
        """Block until nvidia-smi has attached to the GPU and delivered its first sample."""
        t_end = time.perf_counter() + timeout_s
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def mark(self):
        """Called at the start and at the end of the timed region: summary() reports the samples in between (plus one
        on either side: a 100 ms period can leave a short region without a sample of its own)."""
        self.marks.append(len(self.rows))

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if len(self.marks) >= 2:
            rows = self.rows[max(0, self.marks[0] - 1):self.marks[1] + 1] or self.rows
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# per-kernel roofline measurements (CUDA events on the launching stream, L2 flushed between launches)
# ----------------------------------------------------------------------------------------------
def time_kernel(fn, flush, iters=20, warm=3):
    """Median device time of one invocation of `fn` with a cold L2 (CUDA events around a graph replay: an op that is
    several launches -- layout conversion + conv -- would otherwise include the host's launch latency between them)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()  # 256 MiB write: evicts the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts)


def kernel_rooflines(model, peaks, cfg):
    """Achieved algorithmic bytes / FLOPs per launch for the kernels BASELINE.json names, at the configuration's
    cost-volume extent (one pair)."""
    import torch
    from esmstereo_b200 import _lib, ops
    s = cfg["cv"]
    h, w, D, C, G = cfg["H"] // s, cfg["W"] // s, MAXDISP // s, 64, 32
    H, W = cfg["H"], cfg["W"]
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    L = torch.randn(1, C, h, w, generator=g).to(dev)
    R = torch.randn(1, C, h, w, generator=g).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm, fp32_peak, tf32_peak = peaks["hbm_gbs"], peaks["fp32_tflops"], peaks["tf32_tflops"]
    out = {}
    t = time_kernel(lambda: ops.build_gwc_volume(L, R, D, G), flush)
    by = 4.0 * (2 * C * h * w + G * D * h * w)
    out["gwc_volume"] = {"bound": "hbm", "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm,
                         "traffic": None, "us": t * 1e6, "algorithmic_bytes": by}
    # The conv family runs on tcgen05 (split-TF32) where that wins, else on the FP32 pipe.  Whichever path a layer
    # took decides the roofline it is held against: "tensor" = the dense TF32 rate MEASURED in this run
    # (esm_umma_tf32_peak: every SM issuing M128 x N256 x K8 MMAs), "fp32" = the FFMA rate measured on this pool.
    # `achieved` counts ALGORITHMIC FLOPs (2*Cout*Cin*taps*voxels); the split executes 3 (or 2 wider) MMAs per product.
    lib = _lib.lib()
    counters = {"resident tcgen05 (conv_tc.cu)": lib.esm_tc_conv_launches, "streamed tcgen05 (conv_tcg.cu)": lib.esm_tcg_conv_launches,
                "flat TMA-fed tcgen05 (conv_tcf.cu)": lib.esm_tcf_conv_launches}

    def conv_entry(fn, fl, by):
        before = {k: c() for k, c in counters.items()}
        fn()
        path = [k for k, c in counters.items() if c() > before[k]]
        on_tc = bool(path)
        t = time_kernel(fn, flush)
        peak = tf32_peak if on_tc else fp32_peak
        e = {"bound": "tensor" if on_tc else "fp32", "path": (path[0] + ", split-TF32") if on_tc else "fp32 pipe",
             "achieved": fl / t / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": fl / t / 1e12 / peak, "traffic": None,
             "us": t * 1e6, "algorithmic_flops": fl, "algorithmic_bytes": by}
        if on_tc:
            e["executed_mma_tflops"] = 3 * fl / t / 1e12  # hi*hi + lo*hi + hi*lo
        return e

    pc = model.group_stem.packed()
    att = torch.rand(1, G, h, w, generator=g).to(dev) if s == 16 else None
    fp32_only = model.group_stem.fp32_only
    out["gwc_group_stem_fused"] = conv_entry(lambda: ops.conv([L, R], pc, "gelu", gwc_disp=D, in_mul=att, fp32_only=fp32_only),
                                             2.0 * 8 * 32 * 27 * D * h * w, 4.0 * (2 * C + 8 * D) * h * w)
    x8 = torch.randn(1, 8, D, h, w, generator=g).to(dev)
    pa = model.agg.packed()
    out["agg_conv3d_8_8"] = conv_entry(lambda: ops.conv(x8, pa, "gelu", fp32_only=fp32_only), 2.0 * 8 * 8 * 27 * D * h * w, 4.0 * 16 * D * h * w)
    agg = model.aggregation_out
    c1, c2 = agg.conv1[1].conv.weight.shape[0], agg.conv2[1].conv.weight.shape[0]
    d2, h2, w2 = (D - 1) // 2 + 1, (h - 1) // 2 + 1, (w - 1) // 2 + 1
    x1 = torch.randn(1, c1, d2, h2, w2, generator=g).to(dev)
    p1 = agg.conv1[1].packed()
    v1 = d2 * h2 * w2
    out["hourglass_conv3d_level1"] = conv_entry(lambda: ops.conv(x1, p1, "gelu", fp32_only=fp32_only), 2.0 * c1 * c1 * 27 * v1, 4.0 * 2 * c1 * v1)
    d4, h4, w4 = (d2 - 1) // 2 + 1, (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
    x2 = torch.randn(1, c2, d4, h4, w4, generator=g).to(dev)
    p2 = agg.conv2[1].packed()
    v2 = d4 * h4 * w4
    out["hourglass_conv3d_level2"] = conv_entry(lambda: ops.conv(x2, p2, "gelu", fp32_only=fp32_only), 2.0 * c2 * c2 * 27 * v2, 4.0 * 2 * c2 * v2)
    pup = agg.conv2_up.packed()
    out["hourglass_deconv3d_level2"] = conv_entry(lambda: ops.conv(x2, pup, "gelu", out_size=(d2, h2, w2), fp32_only=fp32_only),
                                                  2.0 * c2 * c1 * 64 * v2, 4.0 * (c2 + 8 * c1) * v2)
    # the upsampler's widest 2D layer shape: 32 -> 32 k3 at half resolution (dm4x / ref4x, ESMStereo.py:250-253,191-199);
    # on the kh-in-K tcgen05 kernel of conv_tc.cu since round 2
    up = getattr(model.upsample_module, "dm4x", None)
    if up is not None:
        pu = up[1].packed()
        cu = up[1].conv.weight.shape[0]
        xu = torch.randn(1, cu, H // 2, W // 2, generator=g).to(dev)
        out["upsampler_conv2d_%d_%d_half_res" % (cu, cu)] = conv_entry(lambda: ops.conv(xu, pu, "gelu", fp32_only=up[1].fp32_only), 2.0 * cu * cu * 9 * (H // 2) * (W // 2),
                                                                        4.0 * 2 * cu * (H // 2) * (W // 2))
    cost = torch.randn(1, D, h, w, generator=g).to(dev)
    if s == 4:
        t = time_kernel(lambda: ops.regression_top2(cost), flush)
        name = "regression_top2"
    else:
        t = time_kernel(lambda: ops.disparity_regression(cost, D), flush)
        name = "disparity_regression"
    by = 4.0 * (D + 1) * h * w
    out[name] = {"bound": "hbm", "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm,
                 "traffic": None, "us": t * 1e6, "algorithmic_bytes": by}
    r = model.upsample_module.r
    prev = torch.randn(1, 1, H // r, W // r, generator=g).to(dev)
    res = torch.randn(1, 1, H, W, generator=g).to(dev)
    t = time_kernel(lambda: ops.bilinear_add(prev, res, r, 4.0), flush)
    by = 4.0 * (H * W + H * W / (r * r) + H * W)
    out["bilinear_add_final"] = {"bound": "hbm", "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s",
                                 "frac": by / t / 1e9 / hbm, "traffic": None, "us": t * 1e6, "algorithmic_bytes": by}
    return out


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            m = json.load(f)
        peaks.update(hbm_gbs=m["hbm_gbs"], bf16_tflops=m["bf16_tflops"], source="MEASURED_PEAKS.json")
    # FP32 FMA-pipe peak measured on this pool's B200 with scratch/fma_bench.cu (scalar FFMA, 148 SMs): 72.7 TFLOP/s
    peaks["fp32_tflops"] = 72.7
    # dense TF32 tensor rate: measured live (every SM issuing M128 x N256 x K8 tcgen05.mma back to back)
    from esmstereo_b200 import _lib
    tf = ctypes.c_float(0.0)
    _lib.check(_lib.lib().esm_umma_tf32_peak(16384, ctypes.byref(tf), None), "umma_tf32_peak")
    peaks["tf32_tflops"] = float(tf.value)
    peaks["tf32_source"] = "measured in this run (esm_umma_tf32_peak, csrc/peak.cu)"
    return peaks


# ----------------------------------------------------------------------------------------------
# GPU-eager baseline: the oracle port (same torch ops as the reference) on the same B200 through cuDNN / ATen
# ----------------------------------------------------------------------------------------------
def gpu_eager_baseline(cfg, sd, dev, reps=30):
    """"the real bar on the same box" -- the reference's PyTorch-eager path on this GPU.  The
    oracle port issues the same ATen / cuDNN calls as the reference's modules; protocol of train_sceneflow.py:254-275
    (10 warm-ups, CUDA events around the timed forwards), with cuDNN's TF32 off (fp32-grade, our parity mode) and on
    (PyTorch's default for convolutions)."""
    import torch
    from esmstereo_b200.weights import synthetic_pair
    orc = make_oracle(cfg, sd, device=dev)
    B = cfg["batch"]
    left, right = [t.to(dev) for t in synthetic_pair(B, cfg["H"], cfg["W"], shift=23, seed=0)]
    out = {"protocol": "10 warm-ups + %d timed eager forwards, CUDA events (train_sceneflow.py:254-275)" % reps, "kind": "oracle port on cuda (ATen/cuDNN)"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(10):
                orc(left, right)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                orc(left, right)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"ms_per_step": ms, "pairs_per_s": B / ms * 1e3}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out


def parity_check(cfg, model, orc, dev):
    """End-to-end parity of THIS run's model against the CPU oracle on one pair of the configuration's shape:
    EPE, top-2 index flips (cv4), cost error, confidence error (config E)."""
    import torch
    from esmstereo_b200.weights import synthetic_pair
    left, right = synthetic_pair(1, cfg["H"], cfg["W"], shift=23, seed=7)
    want = orc(left, right)
    model.capture = {}
    with torch.no_grad():
        if cfg["model"] == "ESMStereo_confidence":
            disp, conf = model(left.to(dev), right.to(dev))
        else:
            disp, conf = model(left.to(dev), right.to(dev), train_status=False)[-1], None
    cap = model.capture
    model.capture = None
    out = {"sample": "one %dx%d pair, CPU fp32 oracle vs this run's engines" % (cfg["H"], cfg["W"]),
           "epe_px": float((disp.cpu() - want["disp"]).abs().mean()),
           "cost_rel_err": float((cap["cost"].cpu() - want["cost"]).abs().max() / want["cost"].abs().max())}
    if "top2_idx" in want and "top2_idx" in cap:
        a = torch.sort(cap["top2_idx"].cpu().to(torch.int64), 1)[0]
        b = torch.sort(want["top2_idx"].to(torch.int64), 1)[0]
        out["top2_flips"] = int((a != b).any(1).sum())
        out["pixels"] = int(a.shape[0] * a.shape[2] * a.shape[3])
    if conf is not None:
        out["conf_max_abs_err"] = float((conf.cpu() - want["conf"]).abs().max())
        out["conf_mean_abs_err"] = float((conf.cpu() - want["conf"]).abs().mean())
    return out


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from esmstereo_b200 import GraphedStereo, StereoPipeline, _lib, ops, shard
    from esmstereo_b200.weights import synthetic_pair

    cfg = args.cfg
    H, W, B = cfg["H"], cfg["W"], cfg["batch"]
    conf_model = cfg["model"] == "ESMStereo_confidence"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    extras = rank == 0 and world == 1 and not args.no_extras
    model, sd = build_weights(cfg)

    # ---- CPU baseline FIRST (before any GPU work shares the host with it), then BN calibration for the parity check
    cpu = orc = None
    if rank == 0 and world == 1:
        rate, cores, n, med = cpu_forward_rate(cfg, sd, args.cpu_seconds)
        cpu = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": "%d forwards of ONE %dx%d pair of this workload (median %.0f ms), taken before any GPU work; oracle port of the "
                         "reference forward, torch-CPU fp32" % (n, H, W, med * 1e3)}
    if extras:
        orc = make_oracle(cfg, sd)
        cl, cr = synthetic_pair(1, H, W, shift=23, seed=0)
        sd = orc.calibrate(cl, cr)  # random-init costs are otherwise +-3e-6 and parity is meaningless
        model.load_state_dict(sd)

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = model.to(dev).eval()
    fwd_kw = {} if conf_model else {"train_status": False}

    # rotating pool of distinct input batches larger than L2 (126 MB)
    per_batch = 2 * 3 * H * W * 4 * B
    NPOOL = max(2, min(16, -(-160_000_000 // per_batch)))
    pool = [tuple(t.to(dev) for t in synthetic_pair(B, H, W, shift=23, seed=100 + rank * NPOOL + i)) for i in range(NPOOL)]
    tuned0 = _lib.lib().esm_conv_tuned_calls()
    ops.LAUNCHES = 0
    model(*pool[0], **fwd_kw)
    launches_per_step = ops.LAUNCHES
    graphed = GraphedStereo(model, (B, 3, H, W), **fwd_kw)
    tuned_calls = _lib.lib().esm_conv_tuned_calls() - tuned0
    pick = (lambda out: out[0]) if conf_model else (lambda out: out[-1])

    # the only collective: gather of the per-rank disparities, issued on a side stream every GATHER_EVERY steps so that
    # it overlaps the next replays ("once per sweep, on a side stream")
    GATHER_EVERY = 8
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    gather_ev = torch.cuda.Event() if world > 1 else None
    stash, peer, gather_how = None, None, None
    if world > 1:
        # copy-engine pulls out of symmetric memory (shard.PeerGather) unless ESM_GATHER=nccl or the rendezvous fails
        if os.environ.get("ESM_GATHER", "p2p") != "nccl":
            try:
                peer = shard.PeerGather(GATHER_EVERY * B, (H, W), dev)
                stash = peer.stash
                gather_how = "copy-engine pulls from symmetric memory over NVLink (shard.PeerGather)"
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("[bench] symmetric-memory gather unavailable (%s): NCCL all-gather\n" % (str(e).splitlines()[0][:200],))
                peer = None
        if peer is None:
            stash = torch.empty(GATHER_EVERY * B, H, W, device=dev)
            gather_how = "NCCL all-gather (shard.gather_disparities)"

    def step(i):
        out = pick(graphed(*pool[i % NPOOL]))
        if world > 1:
            j = i % GATHER_EVERY
            if j == 0 and i > 0:
                torch.cuda.current_stream().wait_event(gather_ev)  # the side stream has read the previous sweep
            stash[j * B:(j + 1) * B].copy_(out)
            if j == GATHER_EVERY - 1:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    if peer is not None:
                        peer.gather()
                    else:
                        shard.gather_disparities(stash, world * GATHER_EVERY * B, rank, world, flat=False)  # dataset-ordered view, no re-copy
                    gather_ev.record(side)
        return out

    def barrier():
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    W_, K = max(args.warmup, 3), args.steps
    # The clock sampler (`nvidia-smi -lms`) is started BEFORE the warm-up and the timed region waits for its first
    # sample: a freshly started nvidia-smi attaches to the GPU and stalls the work in flight once, for ~6 ms (measured:
    # K = 200 and K = 2000 steps differed by a constant 5.9 ms) -- inside a short timed region that is 0.1-0.3 ms per step.
    # It keeps sampling through the timed steps; only those samples are reported.
    with ClockSampler(local) as clk:
        clk.wait_first(3.0)
        for i in range(W_):
            step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.profiler.start()  # lets `ncu --profile-from-start off` capture exactly the timed steps
        clk.mark()
        e0.record()
        for i in range(K):
            step(i)
        if world > 1:
            torch.cuda.current_stream().wait_stream(side)
        e1.record()
        barrier()
        clk.mark()
        torch.cuda.profiler.stop()
        t_dev = e0.elapsed_time(e1) * 1e-3

        # ---- end to end through the public API with HOST buffers: every step copies both images from pinned
        # host memory to the device and the result(s) back to pinned host memory, and the host reads each
        # result; `StereoPipeline` overlaps the copies of neighbouring steps with the graph replay.
        hl = [tuple(t.pin_memory() for t in synthetic_pair(B, H, W, shift=23, seed=200 + rank * 4 + i)) for i in range(4)]
        pipe = StereoPipeline(graphed, depth=2, pick=(lambda out: torch.stack(out)) if conf_model else pick)
        checksum = 0.0

        def e2e_run(n):
            nonlocal checksum
            for i in range(n):
                pipe.submit(*hl[i % 4])
                if i >= 1:
                    checksum += float(pipe.result().reshape(-1)[0])  # the caller reads every result
            checksum += float(pipe.result().reshape(-1)[0])

        e2e_run(W_)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        e2e_run(K)
        e1.record()
        barrier()
        t_e2e = max(e0.elapsed_time(e1) * 1e-3, 0.0)
        t_e2e_wall = time.perf_counter() - t0
    times = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    spread = None
    if world > 1:
        allt = [torch.zeros_like(times) for _ in range(world)]
        dist.all_gather(allt, times)
        per_rank = [float(t[0]) for t in allt]
        spread = {"ms_per_step_min": 1e3 * min(per_rank) / K, "ms_per_step_max": 1e3 * max(per_rank) / K}
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(times[0]), float(times[1])

    if rank == 0:
        d2h = H * W * 4 * B * (2 if conf_model else 1)
        line = {
            "metric": cfg["metric"], "value": world * K * B / t_dev, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "pairs_per_step_per_gpu": B,
                       "l2": "%d rotating input batches (%.0f MB) > 126 MB L2" % (NPOOL, NPOOL * per_batch / 1e6), "cuda_graph": True,
                       "parallelism": "pairs sharded by rank",
                       "gather": ("%s of the disparities every %d steps on a side stream" % (gather_how, GATHER_EVERY)) if world > 1 else None},
            "clocks": clk.summary(),
            "e2e": {"value": world * K * B / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": per_batch,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / K, "wall_ms_per_step": 1e3 * t_e2e_wall / K},
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "autotune_calls": int(tuned_calls),
            "cpu_baseline": cpu,
        }
        if spread:
            line["rank_spread"] = spread
        if extras:
            peaks = load_peaks()
            kr = kernel_rooflines(model, peaks, cfg)
            ncu_path = os.path.join(ROOT, "profiles", "traffic.json")
            if args.config == "B" and os.path.exists(ncu_path):  # dram bytes per launch from the committed `ncu --set full` capture
                with open(ncu_path) as f:
                    for k, v in json.load(f).items():
                        if k in kr:
                            kr[k]["traffic"] = v
            dominant = kr["gwc_group_stem_fused"]
            line["roofline"] = {k: dominant[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
            line["roofline_note"] = ("dominant kernel = fused gwc-volume + group_stem conv3d (the largest single launch of the hot path) "
                                     "on the path %s; tensor peak = dense TF32 rate measured in this run (%.0f TFLOP/s: every SM issuing M128 x "
                                     "N256 x K8 tcgen05.mma), fp32 peak = FFMA rate measured on this pool (scratch/fma_bench.cu); HBM peak from "
                                     "%s; `achieved` counts algorithmic FLOPs, the fp32-grade split issues 3 MMAs per product (DESIGN.md); "
                                     "`traffic` = dram bytes per launch from the committed ncu --set full capture (profiles/)"
                                     % (dominant.get("path", "?"), peaks["tf32_tflops"], peaks["source"]))
            line["peaks"] = peaks
            line["kernels"] = kr
            line["cost_volume_hbm_gbs"] = kr["gwc_volume"]["achieved"]
            line["parity"] = parity_check(cfg, model, orc, dev)
            line["gpu_eager_baseline"] = gpu_eager_baseline(cfg, sd, dev)
        if args.breakdown:
            line["breakdown_ms"] = op_breakdown(model, pool[0], fwd_kw)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def op_breakdown(model, pair, fwd_kw):
    """Eager-mode per-operator device times (CUDA events), for profiling only."""
    import torch
    from esmstereo_b200 import ops
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.PROFILE = []
    torch.cuda.synchronize()
    torch.cuda._sleep(60_000_000)  # keep the GPU busy (~30 ms) so the host runs ahead: event gaps are then pure device time
    e0.record()
    model(*pair, **fwd_kw)
    e1.record()
    torch.cuda.synchronize()
    rows = [(lbl, a.elapsed_time(b)) for lbl, a, b in ops.PROFILE]
    ops.PROFILE = None
    total_ops = sum(t for _, t in rows)
    sys.stderr.write("---- per-operator device time (eager forward queued behind a busy GPU: pure device time) ----\n")
    for lbl, t in rows:
        sys.stderr.write("%8.1f us  %s\n" % (t * 1e3, lbl))
    sys.stderr.write("hot-path ops total %.3f ms; forward wall (eager) %.3f ms\n" % (total_ops, e0.elapsed_time(e1)))
    agg = {}
    for lbl, t in rows:
        key = lbl.split(" in ")[0]
        agg[key] = agg.get(key, 0.0) + t
    return {"hot_path_ops_ms": total_ops, "top": sorted(((round(v, 4), k) for k, v in agg.items()), reverse=True)[:12]}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
