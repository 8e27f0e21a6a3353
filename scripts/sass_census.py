"""SASS opcode census of libesm_b200.so: which kernels carry the Blackwell instructions (B200_PROFILING.md):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG = TMA tensor load, UBLKCP = TMA bulk copy,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, FFMA2 = packed fp32 FMA.  Writes profiles/r02_sass_census.txt.

    python scripts/sass_census.py            # no GPU needed: cuobjdump -sass on the built library
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "esmstereo_b200", "csrc", "libesm_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "FFMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            op = m.group(1)
            per[cur]["total"] += 1
            for o in OPS:
                if op == o or (o == "FFMA" and op == "FFMA"):
                    per[cur][o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, c), dn in zip(per.items(), demangle):
        if not any(c[o] for o in OPS[:7]):
            continue
        short = re.sub(r"\(.*", "", dn).replace("void esm::", "").replace("esm::", "")
        rows.append((short, c))
    path = os.path.join(ROOT, "profiles", "r02_sass_census.txt")
    with open(path, "w") as f:
        f.write("# cuobjdump -sass esmstereo_b200/csrc/libesm_b200.so (sm_100a): instruction counts per kernel; only kernels with\n"
                "# tensor-core / TMA / mbarrier instructions are listed.  scripts/sass_census.py\n")
        f.write("%-58s %7s" % ("kernel", "instrs") + "".join(" %8s" % o for o in OPS) + "\n")
        tot = collections.Counter()
        for short, c in rows:
            f.write("%-58s %7d" % (short[:58], c["total"]) + "".join(" %8d" % c[o] for o in OPS) + "\n")
            tot.update(c)
        f.write("%-58s %7d" % ("TOTAL (%d kernels)" % len(rows), tot["total"]) + "".join(" %8d" % tot[o] for o in OPS) + "\n")
    sys.stdout.write(open(path).read())


if __name__ == "__main__":
    main()
