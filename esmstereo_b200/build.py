"""Build libesm_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m esmstereo_b200.build [--force]

The library is built next to its sources (esmstereo_b200/csrc/libesm_b200.so) so that it travels to
the GPU box with the repo snapshot; nothing is JIT-compiled at import time.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libesm_b200.so")
SOURCES = ["api.cu", "conv.cu", "conv_k1.cu", "conv_k2.cu", "conv_k3.cu", "conv_k3s2.cu", "conv_k5.cu", "conv_tc.cu", "conv_tcg.cu", "conv_tcf.cu", "conv_pw.cu", "conv_stem3.cu", "volume.cu", "regress.cu",
           "mixer.cu", "conf.cu", "prepost.cu", "peak.cu", "backbone.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> str:
    nvcc = _nvcc()
    prof = os.environ.get("ESM_TC_PROFILE") == "1"  # diagnostics build: its own objects and library, next to the product's
    ab = os.environ.get("ESM_AB_DEFS", "").split()  # A/B build: extra -D flags, its own objects and library (load it with ESM_LIB)
    tag = os.environ.get("ESM_AB_TAG", "ab")
    objdir = os.path.join(CSRC, "build_prof" if prof else "build_" + tag if ab else "build")
    LIB = os.path.join(CSRC, "libesm_b200_prof.so" if prof else "libesm_b200_%s.so" % tag if ab else "libesm_b200.so")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "conv_kernel.cuh"), os.path.join(CSRC, "conv_tc.cuh"), os.path.join(CSRC, "tc_common.cuh"),
               os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "esm_b200.h")]
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if os.environ.get("ESM_TC_PROFILE") == "1":  # role timers of the tcgen05 conv kernel (conv_tc.cu), diagnostics only
        flags = flags + ["-DTC_PROFILE"]
    flags = flags + ab

    # objects built with other flags (e.g. the ESM_TC_PROFILE role timers) are stale too
    stamp = os.path.join(objdir, "flags.txt")
    flag_text = " ".join(flags)
    if not os.path.exists(stamp) or open(stamp).read() != flag_text:
        force = True
        with open(stamp, "w") as f:
            f.write(flag_text)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + flags + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
