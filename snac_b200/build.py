"""In-tree build of libdmp.so (hand-written sm_100a kernels + the C ABI of include/dmp.h).

    python snac_b200/build.py          # or: from snac_b200.build import build_lib; build_lib()

nvcc cross-compiles for sm_100a without a GPU; the resulting snac_b200/libdmp.so is git-ignored but
travels with gpurun snapshots.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libdmp.so")
SOURCES = ["dmp_api.cu", "dmp_1d.cu", "dmp_2d.cu", "dmp_3d.cu", "dmp_3d_roll.cu", "dmp_3d_step.cu", "dmp_stages.cu", "dmp_plangen.cu"]
HEADERS = [os.path.join(CSRC, h) for h in ("dmp_common.cuh", "dmp_3d_bulk.cuh")] + [os.path.join(ROOT, "include", "dmp.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
] + os.environ.get("SNAC_B200_NVCC_FLAGS", "").split()      # build-time experiments only (e.g. -DDMP3_EARLY_ROWS=1)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(PKG, "build", os.path.basename(s) + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: %s" % " ".join(cmd))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
