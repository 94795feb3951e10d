"""Builds libpn2_b200.so in-tree with nvcc for sm_100a (no torch, no pybind).

    python -m pointnet2.scannet_b200.build [--force] [--verbose]

The library has a plain C ABI (include/pn2_b200.h) and links only libcudart, so
one nvcc command is the whole build.  Objects are cached per source under
csrc/build/ and rebuilt when the source or a header is newer.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(PKG))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libpn2_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]
# Index parity needs un-contracted fp32 arithmetic.  The distance code uses
# __fmul_rn/__fadd_rn explicitly; -fmad=false on the op kernels is the belt to
# those braces.  The tensor-core MLP translation unit may contract.
NO_FMA = {"fps.cu", "ball_query.cu", "three_nn.cu", "group.cu", "interpolate.cu"}


def _newest_header():
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(src), _newest_header())):
        return obj, ""
    cmd = [NVCC] + NVCC_FLAGS + (["-fmad=false"] if os.path.basename(src) in NO_FMA else []) \
        + ["-Xptxas", "-v", "-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd))
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    if os.path.basename(src) in NO_FMA:
        try:
            _assert_no_contraction(obj)
        except RuntimeError:
            os.remove(obj)  # never leave an object behind that a later incremental build would link
            raise
    return obj, p.stderr


def _assert_no_contraction(obj):
    """Index parity depends on un-contracted fp32 distances: the only fused multiply-adds allowed
    in these objects are the packed squares written as fma(d, d, +0) (addend RZ)."""
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    bad = [l for l in sass.splitlines()
           if (" FFMA " in l or " FFMA2 " in l or " FFMA." in l) and "RZ" not in l.split(";")[0].split(",")[-1]]
    if bad:
        raise RuntimeError("contracted FMA in %s:\n%s" % (obj, "\n".join(bad[:5])))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
        if verbose:
            print(log)
    if (force or not os.path.exists(LIB)
            or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs)):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared",
               "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
