"""Turn the ncu captures in gpurun_out/ into the committed text summaries under profiles/.

    python scripts/summarize_profiles.py r01
"""
import collections
import csv
import json
import os
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = "profiles"
os.makedirs(OUT, exist_ok=True)


def launches():
    rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
    hdr, data = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [r[ix["Kernel Name"]] for r in data]
    half = len(data) // 2  # the capture holds exactly the 2 timed (graph-replayed) steps
    step = data[half:]
    total = sum(float(r[ix["Metric Value"]]) for r in step)
    agg = collections.OrderedDict()
    for r in step:
        n = r[ix["Kernel Name"]].replace("esm::", "")[:64]
        v = agg.setdefault(n, [0, 0.0])
        v[0] += 1
        v[1] += float(r[ix["Metric Value"]])
    with open(os.path.join(OUT, "%s_launches_bench_step.txt" % TAG), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off  python bench.py --steps 2 --warmup 3 --cpu-seconds 1\n")
        f.write("# second of the two timed, graph-replayed steps; cold-cache, serialised:\n")
        f.write("# compare SHARES, not absolutes.  %d launches, sum %.3f ms\n" % (len(step), total / 1e6))
        f.write("# share%%   total_us  launches  kernel\n")
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%6.2f  %9.1f  %5d  %s\n" % (100 * v / total, v / 1e3, c, n))
        f.write("\n# in launch order: us, grid, block, kernel\n")
        for r in step:
            f.write("%9.1f  %-12s %-10s %s\n" % (float(r[ix["Metric Value"]]) / 1e3, r[ix["Grid Size"]].replace(" ", ""),
                                                 r[ix["Block Size"]].replace(" ", ""), r[ix["Kernel Name"]].replace("esm::", "")[:80]))
    print("launch list:", len(step), "launches", total / 1e6, "ms")


METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
           "sm__inst_executed_pipe_uniform.sum",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def full():
    rep = "gpurun_out/prof_kernels.ncu-rep"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    label = {"gwc_volume_kernel": "gwc_volume", "regression_top2_kernel": "regression_top2", "bilinear_add_kernel": "bilinear_add_final"}
    with open(os.path.join(OUT, "%s_ncu_full_kernels.txt" % TAG), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on --profile-from-start off  python scripts/prof_conv.py\n")
        f.write("# (KITTI shapes: h x w = 96x312, D=48; one launch per kernel, tuned plans)\n")
        for r in data:
            name = r[ix["Kernel Name"]]
            f.write("\n== %s\n" % name)
            for m in METRICS:
                if m in ix:
                    f.write("   %-82s %s %s\n" % (m, r[ix[m]], units[ix[m]]))
            def val(m):
                v = float(r[ix[m]].replace(",", ""))
                u = units[ix[m]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            byts = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            key = None
            for k, v in label.items():
                if k in name:
                    key = v
            if "tc_conv_kernel<8, 3, 3, 1" in name or ("conv_kernel" in name and "tc_" not in name and ", 1, 1, 0" in name):
                key = "gwc_group_stem_fused"
            if "tc_conv_kernel<8, 3, 3, 0" in name:
                key = "agg_conv3d_8_8"
            if "tc_conv_kernel<24, 1, 3, 0" in name:
                key = "hourglass_conv3d_24_24"
            if "tc_conv_kernel<24, 1, 3, 0" in name:
                key = "hourglass_conv3d_level1"
            if "tck_conv_kernel" in name:
                key = "upsampler_conv2d_32_32_half_res"
            if "tcf_conv_kernel" in name and "hourglass_conv3d_level2" not in traffic:
                key = "hourglass_conv3d_level2"      # prof_conv.py runs the 40 -> 40 conv before the 40 -> 24 transposed conv
            elif "tcf_conv_kernel" in name:
                key = "hourglass_deconv3d_level2"
            if key:
                traffic[key] = byts
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print("traffic", traffic)
    # source-level hot spots of the dominant kernel
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:conv_kernel|tck_conv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif cur is not None and r:
            if r[0] == "Address":
                cur["hdr"] = r
            else:
                cur["rows"].append(r)
    seen = set()
    with open(os.path.join(OUT, "%s_ncu_conv_stalls.txt" % TAG), "w") as f:
        for sec in secs:
            if sec["name"] in seen:
                continue
            seen.add(sec["name"])
            hdr = sec["hdr"]
            jx = {h: i for i, h in enumerate(hdr)}
            cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot, byop, iop, T = collections.Counter(), collections.Counter(), collections.Counter(), 0
            for r in sec["rows"]:
                try:
                    s, n = int(r[jx["# Samples"]]), int(r[jx["Instructions Executed"]])
                except ValueError:
                    continue
                T += s
                toks = r[jx["Source"]].split()
                op = toks[0] if toks else "?"
                if op.startswith("@") and len(toks) > 1:
                    op = toks[1]
                op = op.split(".")[0]
                byop[op] += s
                iop[op] += n
                for h in cols:
                    if r[jx[h]].isdigit():
                        tot[h] += int(r[jx[h]])
            f.write("== %s  (%d SASS instructions, %d warp samples)\n" % (sec["name"], len(sec["rows"]), T))
            f.write("   stall reasons (%% of samples): %s\n" % ", ".join("%s %.1f" % (k[6:], 100 * v / T) for k, v in tot.most_common(10)))
            f.write("   opcodes by samples (%% samples / M warp-instructions): %s\n\n" % ", ".join(
                "%s %.1f/%.1f" % (op, 100 * s / T, iop[op] / 1e6) for op, s in byop.most_common(12)))


if __name__ == "__main__":
    launches()
    full()
