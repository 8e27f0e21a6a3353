"""Summarise an ncu report: per kernel the headline metrics and the hot SASS regions (by executed
warp-instructions and stall samples) with the CUDA source lines they map to.

    python scripts/ncu_hot.py gpurun_out/prof.ncu-rep
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.sum"]
for r in rows[2:]:
    print("; ".join("%s=%s" % (k.split(".")[0][-28:], r[hdr.index(k)]) for k in KEYS if k in hdr))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
if "Kernel Name" not in src:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] in ("Address", "#"):
        cur["hdr"] = r
    elif cur is not None and len(r) > 10:
        cur["rows"].append(r)
seen = set()
for k in kern:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    h = k["hdr"]
    ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[ie]) for r in k["rows"])
    ts = sum(int(r[isamp]) for r in k["rows"]) or 1
    print("\n==", k["name"][:80], "warp-instr", tot, "samples", ts, "cols:", h[:3])
    op = collections.Counter()
    for r in k["rows"]:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
        op[m.group(2) if m else "?"] += int(r[ie])
    print("  opcodes %:", " ".join("%s %.1f" % (o, 100 * c / tot) for o, c in op.most_common(22)))
