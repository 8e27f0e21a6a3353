"""Benchmark of the ESMStereo hot path (BASELINE.json: stereo pairs/sec @384x1248 maxdisp192; cost-volume
HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--breakdown]

One step = one forward of ESMStereo-L (cv4, gwc, maxdisp 192) over one synthetic KITTI-shaped stereo
pair (BASELINE.json configs[1]) per GPU; pairs shard by rank with no data-path collective (weak
scaling), and the per-step disparities are all-gathered over NCCL as the reference-side "gather of
outputs".  Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port instead
(the Python reference cannot travel to the GPU box).
"""
import argparse
import contextlib
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

H, W, MAXDISP = 384, 1248, 192
WORKLOAD = "ESMStereo-L (cv_scale 4, gwc, efficientnet_b2 stand-in backbone) 384x1248 pair, batch 1 per GPU, maxdisp 192, fp32"
METRIC = "stereo pairs/sec @384x1248 maxdisp192"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--breakdown", action="store_true", help="also print a per-operator timing table to stderr")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    return ap.parse_args()


def build_weights(seed=0):
    import torch  # noqa: F401
    from esmstereo_b200 import __models__
    from esmstereo_b200.weights import fill_deterministic
    with contextlib.redirect_stdout(io.StringIO()):
        model = __models__["ESMStereo"](MAXDISP, True, False, "efficientnet_b2", 4)
    sd = fill_deterministic(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    return model, sd


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference forward, all host threads
# ----------------------------------------------------------------------------------------------
def cpu_forward_rate(sd, budget_s, min_iters=2):
    import torch
    from esmstereo_b200.weights import synthetic_pair
    from oracle.esm_oracle import EsmOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = EsmOracle(sd, MAXDISP, True, False, "efficientnet_b2", 4)
    left, right = synthetic_pair(1, H, W, shift=23, seed=0)
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
    _, sd = build_weights()
    import torch
    from esmstereo_b200.weights import synthetic_pair
    from oracle.esm_oracle import EsmOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = EsmOracle(sd, MAXDISP, True, False, "efficientnet_b2", 4)
    left, right = synthetic_pair(1, H, W, shift=23, seed=0)
    steps = max(1, min(args.steps, 20))      # bounded sample: a CPU forward is ~1 s
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        orc(left, right)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc(left, right)
    dt = time.perf_counter() - t0
    value = steps / dt
    sample = "%d forwards of one 384x1248 pair (of the %d steps asked), torch-CPU fp32, %d threads" % (steps, args.steps, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "arm": "CPU oracle port of the reference forward"},
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
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()  # 256 MiB write: evicts the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sum(ts) / len(ts)


def kernel_rooflines(model, peaks):
    """Achieved algorithmic bytes / FLOPs per launch for the kernels BASELINE.json names."""
    import torch
    from esmstereo_b200 import ops
    h, w, D, C, G = H // 4, W // 4, MAXDISP // 4, 64, 32
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    L = torch.randn(1, C, h, w, generator=g).to(dev)
    R = torch.randn(1, C, h, w, generator=g).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm, fp32_peak = peaks["hbm_gbs"], peaks["fp32_tflops"]
    out = {}
    # cost volume, standalone (K1): 4*(2*C*h*w + G*D*h*w) bytes
    t = time_kernel(lambda: ops.build_gwc_volume(L, R, D, G), flush)
    by = 4.0 * (2 * C * h * w + G * D * h * w)
    out["gwc_volume"] = {"bound": "hbm", "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm,
                         "traffic": None, "us": t * 1e6, "algorithmic_bytes": by}
    # The conv family runs on tcgen05 (split-TF32, conv_tc.cu) where that wins the on-device timing, else on
    # the FP32 pipe.  Whichever path a layer took decides the roofline it is held against: "tensor" = dense
    # TF32 peak = half of the measured bf16 burst peak (kind::tf32 runs at half the bf16 rate; MEASURED_PEAKS
    # has no separate TF32 figure), "fp32" = the FFMA rate measured on this pool.  `achieved` counts the
    # ALGORITHMIC FLOPs (2*Cout*Cin*taps*voxels); the split executes 3 MMAs per product, reported separately.
    from esmstereo_b200 import _lib
    tc_count, tcg_count = _lib.lib().esm_tc_conv_launches, _lib.lib().esm_tcg_conv_launches
    tf32_peak = peaks["bf16_tflops"] / 2.0

    def conv_entry(fn, fl, by):
        n0, g0 = tc_count(), tcg_count()
        fn()
        on_tcg = tcg_count() > g0
        on_tc = on_tcg or tc_count() > n0
        t = time_kernel(fn, flush)
        peak = tf32_peak if on_tc else fp32_peak
        path = "tcgen05 split-TF32, streamed weights (conv_tcg.cu)" if on_tcg else "tcgen05 split-TF32, resident weights (conv_tc.cu)"
        e = {"bound": "tensor" if on_tc else "fp32", "path": path if on_tc else "fp32 pipe",
             "achieved": fl / t / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": fl / t / 1e12 / peak, "traffic": None,
             "us": t * 1e6, "algorithmic_flops": fl, "algorithmic_bytes": by}
        if on_tc:
            e["executed_mma_tflops"] = 3 * fl / t / 1e12  # hi*hi + lo*hi + hi*lo
        return e

    # fused volume + group_stem (K1 fused into K2): 2*8*32*27*voxels FLOPs, 4*(2*C + 8*D)*h*w bytes
    pc = model.group_stem.packed()
    out["gwc_group_stem_fused"] = conv_entry(lambda: ops.conv([L, R], pc, "gelu", gwc_disp=D), 2.0 * 8 * 32 * 27 * D * h * w,
                                             4.0 * (2 * C + 8 * D) * h * w)
    # agg (a4): 8 -> 8 k3 at full cost-volume resolution; aggregation.conv1.1 (a5): 24 -> 24 k3 at half resolution
    x8 = torch.randn(1, 8, D, h, w, generator=g).to(dev)
    pa = model.agg.packed()
    out["agg_conv3d_8_8"] = conv_entry(lambda: ops.conv(x8, pa, "gelu"), 2.0 * 8 * 8 * 27 * D * h * w, 4.0 * 16 * D * h * w)
    x24 = torch.randn(1, 24, D // 2, h // 2, w // 2, generator=g).to(dev)
    p24 = model.aggregation_out.conv1[1].packed()
    out["hourglass_conv3d_24_24"] = conv_entry(lambda: ops.conv(x24, p24, "gelu"), 2.0 * 24 * 24 * 27 * (D // 2) * (h // 2) * (w // 2),
                                               4.0 * 48 * (D // 2) * (h // 2) * (w // 2))
    # the wide hourglass level (a5): 40 -> 40 k3 at quarter resolution, and its ConvTranspose3d k4 s2 40 -> 24
    x40 = torch.randn(1, 40, D // 4, h // 4, w // 4, generator=g).to(dev)
    p40 = model.aggregation_out.conv2[1].packed()
    v40 = (D // 4) * (h // 4) * (w // 4)
    out["hourglass_conv3d_40_40"] = conv_entry(lambda: ops.conv(x40, p40, "gelu"), 2.0 * 40 * 40 * 27 * v40, 4.0 * 80 * v40)
    pup = model.aggregation_out.conv2_up.packed()
    out["hourglass_deconv3d_40_24"] = conv_entry(lambda: ops.conv(x40, pup, "gelu"), 2.0 * 40 * 24 * 64 * v40, 4.0 * (40 + 8 * 24) * v40)
    # regression (K4): 4*(D+1)*h*w bytes
    cost = torch.randn(1, D, h, w, generator=g).to(dev)
    t = time_kernel(lambda: ops.regression_top2(cost), flush)
    by = 4.0 * (D + 1) * h * w
    out["regression_top2"] = {"bound": "hbm", "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm,
                              "traffic": None, "us": t * 1e6, "algorithmic_bytes": by}
    # final assembly (a11): read residual HxW + previous H/2xW/2, write HxW
    prev = torch.randn(1, 1, H // 2, W // 2, generator=g).to(dev)
    res = torch.randn(1, 1, H, W, generator=g).to(dev)
    t = time_kernel(lambda: ops.bilinear_add(prev, res, 2, 4.0), flush)
    by = 4.0 * (H * W + H * W / 4 + H * W)
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
    return peaks


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from esmstereo_b200 import GraphedStereo, ops  # noqa: F401
    from esmstereo_b200.weights import synthetic_pair

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, sd = build_weights()
    model = model.to(dev).eval()

    # rotating pool of distinct input pairs larger than L2 (16 x 11.5 MB = 184 MB > 126 MB)
    NPOOL = 16
    pool = [tuple(t.to(dev) for t in synthetic_pair(1, H, W, shift=23, seed=100 + rank * NPOOL + i)) for i in range(NPOOL)]
    ops.LAUNCHES = 0
    model(*pool[0], train_status=False)
    launches_per_step = ops.LAUNCHES
    graphed = GraphedStereo(model, (1, 3, H, W), train_status=False)
    gathered = torch.empty(world, H, W, device=dev) if world > 1 else None

    def step(i):
        out = graphed(*pool[i % NPOOL])[-1]
        if world > 1:  # the only collective: gather of the per-rank disparities (no data-path exchange)
            dist.all_gather_into_tensor(gathered, out[0])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W_, K = max(args.warmup, 3), args.steps
    for i in range(W_):
        step(i)
    barrier()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.profiler.start()  # lets `ncu --profile-from-start off` capture exactly the timed steps
        e0.record()
        for i in range(K):
            step(i)
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
        t_dev = e0.elapsed_time(e1) * 1e-3

        # ---- end to end through the public API with HOST buffers: every step copies both images from pinned
        # host memory to the device and the disparity back to pinned host memory, and the host reads each
        # result; `StereoPipeline` overlaps the copies of neighbouring steps with the graph replay.
        from esmstereo_b200 import StereoPipeline
        hl = [tuple(t.pin_memory() for t in synthetic_pair(1, H, W, shift=23, seed=200 + rank * 4 + i)) for i in range(4)]
        pipe = StereoPipeline(graphed, depth=2)
        checksum = 0.0

        def e2e_run(n):
            nonlocal checksum
            for i in range(n):
                pipe.submit(*hl[i % 4])
                if i >= 1:
                    checksum += float(pipe.result()[0, 0, 0])  # the caller reads every result
            checksum += float(pipe.result()[0, 0, 0])

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
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(times[0]), float(times[1])

    if rank == 0:
        peaks = load_peaks()
        kr = kernel_rooflines(model, peaks)
        ncu_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(ncu_path):  # dram bytes per launch from the committed `ncu --set full` capture
            with open(ncu_path) as f:
                for k, v in json.load(f).items():
                    if k in kr:
                        kr[k]["traffic"] = v
        breakdown = None
        if args.breakdown:
            breakdown = op_breakdown(model, pool[0])
        cpu = None
        if world == 1:
            rate, cores, n, med = cpu_forward_rate(sd, args.cpu_seconds)
            cpu = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
                   "sample": "%d forwards of the same 384x1248 pair workload (median %.0f ms), oracle port of the reference "
                             "forward, torch-CPU fp32" % (n, med * 1e3)}
        dominant = kr["gwc_group_stem_fused"]
        line = {
            "metric": METRIC, "value": world * K / t_dev, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": 1, "l2": "%d rotating input pairs (%.0f MB) > 126 MB L2"
                       % (NPOOL, NPOOL * 2 * 3 * H * W * 4 / 1e6), "cuda_graph": True, "parallelism": "pairs sharded by rank"},
            "clocks": clk.summary(),
            "e2e": {"value": world * K / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": 2 * 3 * H * W * 4,
                    "d2h_bytes_per_step": H * W * 4, "ms_per_step": 1e3 * t_e2e / K, "wall_ms_per_step": 1e3 * t_e2e_wall / K},
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "roofline": {k: dominant[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")},
            "roofline_note": "dominant kernel = fused gwc-volume + group_stem conv3d (19.9 GFLOP/launch, the largest single "
                             "launch of the step) on the path the autotuner chose (%s); tensor peak = dense TF32 = measured bf16 "
                             "burst / 2, fp32 peak = FFMA rate measured on this pool (scratch/fma_bench.cu); peaks from %s; "
                             "`achieved` counts algorithmic FLOPs, the fp32-grade split issues 3 MMAs per product (DESIGN.md)"
                             % (dominant.get("path", "?"), peaks["source"]),
            "kernels": kr,
            "cost_volume_hbm_gbs": kr["gwc_volume"]["achieved"],
            "cpu_baseline": cpu,
        }
        if breakdown is not None:
            line["breakdown_ms"] = breakdown
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def op_breakdown(model, pair):
    """Eager-mode per-operator device times (CUDA events), for profiling only."""
    import torch
    from esmstereo_b200 import ops
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.PROFILE = []
    torch.cuda.synchronize()
    torch.cuda._sleep(60_000_000)  # keep the GPU busy (~30 ms) so the host runs ahead: event gaps are then pure device time
    e0.record()
    model(*pair, train_status=False)
    e1.record()
    torch.cuda.synchronize()
    rows = [(lbl, a.elapsed_time(b)) for lbl, a, b in ops.PROFILE]
    ops.PROFILE = None
    total_ops = sum(t for _, t in rows)
    sys.stderr.write("---- per-operator device time (eager forward queued behind a busy GPU: pure device time) ----\n")
    import re
    for lbl, t in rows:
        m = re.match(r"(de)?conv(\d)d (\d+)->(\d+) k(\d)(\+gwc)? s(\d) in (\d+)x(\d+)x(\d+)", lbl)
        extra = ""
        if m:
            dec, nd, cin, cout, k, _g, st, D, Hh, Ww = m.groups()
            nd, cin, cout, k, st, D, Hh, Ww = int(nd), int(cin), int(cout), int(k), int(st), int(D), int(Hh), int(Ww)
            taps = k ** nd
            vox = D * Hh * Ww  # input voxels
            if dec:
                fl = 2.0 * cin * cout * taps * vox
            else:
                out_vox = vox / (st ** nd) if st > 1 else vox
                fl = 2.0 * cin * cout * taps * out_vox
            fl *= pair[0].shape[0] * (2 if lbl.startswith("conv2d 3->") or False else 1)
            extra = "  %6.2f GFLOP %5.1f TFLOP/s" % (fl / 1e9, fl / (t * 1e-3) / 1e12)
        sys.stderr.write("%8.1f us  %s%s\n" % (t * 1e3, lbl, extra))
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
