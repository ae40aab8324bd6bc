"""Builds shapeclipper_b200/libsc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m shapeclipper_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
import glob
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_obj")
LIB = os.path.join(PKG, "libsc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = (["-DSC_TC_TRACE"] if os.environ.get("SC_TC_TRACE") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stamp(src):
    h = hashlib.sha1()
    for p in [src] + sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs, relink = [], force or not os.path.isfile(LIB)
    procs = []
    for src in srcs:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp(src)
        objs.append(obj)
        if not force and os.path.isfile(obj) and os.path.isfile(stamp_file) and open(stamp_file).read() == stamp:
            continue
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((subprocess.Popen(cmd), stamp_file, stamp, src))
        relink = True
    for p, stamp_file, stamp, src in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on %s" % src)
        with open(stamp_file, "w") as f:
            f.write(stamp)
    if relink:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
