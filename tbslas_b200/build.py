"""Build libtbslas_b200.so in-tree with nvcc for sm_100a (no torch dependency).

    python -m tbslas_b200.build [--force] [--jobs N]

Objects go to tbslas_b200/csrc/build/, the library to tbslas_b200/libtbslas_b200.so
(git-ignored; it travels to the GPU box with the gpurun snapshot).  The CUDA runtime is
linked statically so the library loads on a CPU-only host too (symbol checks); NCCL is
dlopen'ed at comm_init time.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libtbslas_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-ffp-contract=off", "-ccbin", "/usr/bin/g++",
         "-I", CSRC, "-I", os.path.join(ROOT, "include")]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool, objdir: str = OBJ, extra=()) -> str:
    obj = os.path.join(objdir, os.path.basename(src) + ".o")
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _stale(obj, [src] + hdrs):
        cmd = [NVCC] + FLAGS + list(extra) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr))
    return obj


def build(force: bool = False, jobs: int = 0, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant/defines: an A/B build of the same sources with extra -D flags, written to
    tbslas_b200/variants/libtbslas_b200_<variant>.so (load it with TBSLAS_B200_LIB=<path>);
    kernel experiments only -- the product is the default build."""
    objdir, lib = OBJ, LIB
    if variant:
        objdir = os.path.join(CSRC, "build_" + variant)
        os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
        lib = os.path.join(HERE, "variants", "libtbslas_b200_%s.so" % variant)
    os.makedirs(objdir, exist_ok=True)
    extra = ["-D" + d for d in defines]
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = jobs or min(len(srcs), os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, objdir, extra), srcs))
    LIB_OUT = lib
    if force or _stale(LIB_OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_OUT] + objs + ["-cudart", "static", "-ccbin", "/usr/bin/g++",
                                                      "-Xlinker", "--no-undefined", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr))
    if verbose:
        print("built", LIB_OUT)
    return LIB_OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=0)
    ap.add_argument("--variant", default="")
    ap.add_argument("-D", dest="defines", action="append", default=[])
    a = ap.parse_args()
    build(a.force, a.jobs, verbose=True, variant=a.variant, defines=a.defines)
    sys.exit(0)
