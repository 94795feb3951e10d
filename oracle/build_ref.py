"""Compile the reference's own CUDA extension, unmodified, for sm_100a.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Sources are compiled where
they lie under /root/reference/pointnet2/_ext_src (never copied); only build
products land in oracle/_ref/ (git-ignored, but shipped to the GPU box):

  oracle/_ref/nofma/pn2_ref_nofma.so   -O3 -fmad=false   NORMATIVE parity oracle
  oracle/_ref/stock/pn2_ref_stock.so   -O3               the reference's own flags
                                                         (pointnet2/setup.py:30-33):
                                                         performance baseline

The reference's setup.py cannot be used as is: it pins TORCH_CUDA_ARCH_LIST to
3.7+PTX..7.5 (pointnet2/setup.py:17) and nvcc 12.9 rejects compute_37.  The
same source list (setup.py:9-13) is handed to torch.utils.cpp_extension.load
with the arch overridden to 10.0a.

stage_python() additionally places VERBATIM copies of the reference's own model code
(pointnet2/{pointnet2_semseg,pointnet2_modules,pointnet2_utils,pytorch_utils}.py and
lib/loss.py) into oracle/_ref/py/ -- git-ignored like the rest of oracle/_ref, shipped to the
GPU box -- so that `bench.py --impl reference` runs the unmodified reference end to end
(its Python over its kernels) where /root/reference does not exist.

Both modules expose the nine functions of bindings.cpp:6-19 and need a GPU to
run ("CPU not supported", e.g. ball_query.cpp:27-29).
"""
import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pointnet2/_ext_src"
OUT = os.path.join(HERE, "_ref")
VARIANTS = {
    "nofma": ["-O3", "-fmad=false"],
    "stock": ["-O3"],
}


def so_path(variant):
    return os.path.join(OUT, variant, "pn2_ref_%s.so" % variant)


def available(variant="nofma"):
    return os.path.exists(so_path(variant))


def build(variant, verbose=False):
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference sources not present at %s" % REF_SRC)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    bdir = os.path.join(OUT, variant)
    os.makedirs(bdir, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp"))
                     + glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    load(
        name="pn2_ref_%s" % variant,
        sources=sources,
        extra_include_paths=[os.path.join(REF_SRC, "include")],
        extra_cflags=["-O2"],
        extra_cuda_cflags=VARIANTS[variant] + ["-lineinfo"],
        build_directory=bdir,
        verbose=verbose,
        is_python_module=True,
    )
    return so_path(variant)


def python_dir():
    return os.path.join(OUT, "py")


REF_PY = ["pointnet2_semseg.py", "pointnet2_modules.py", "pointnet2_utils.py", "pytorch_utils.py"]


def stage_python():
    """Verbatim copies of the reference's model code and loss into oracle/_ref/py (see module docstring)."""
    import shutil
    ref_root = os.path.dirname(os.path.dirname(REF_SRC))
    dst = python_dir()
    os.makedirs(os.path.join(dst, "solver", "lib"), exist_ok=True)
    for f in REF_PY:
        shutil.copyfile(os.path.join(ref_root, "pointnet2", f), os.path.join(dst, f))
    shutil.copyfile(os.path.join(ref_root, "lib", "loss.py"), os.path.join(dst, "solver", "lib", "loss.py"))
    open(os.path.join(dst, "solver", "lib", "__init__.py"), "w").close()
    return dst


def build_all(verbose=False):
    for v in VARIANTS:
        if not available(v):
            build(v, verbose=verbose)
    stage_python()


def load_ref(variant="nofma"):
    """Import a previously built variant (works on the GPU box: no sources needed)."""
    path = so_path(variant)
    if not os.path.exists(path):
        raise ImportError("oracle/_ref/%s not built (run oracle/build_ref.py where "
                          "/root/reference exists)" % variant)
    import torch  # noqa: F401  (libtorch must be loaded first)
    name = "pn2_ref_%s" % variant
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == "__main__":
    for v in (sys.argv[1:] or list(VARIANTS)):
        print("building", v, "->", build(v, verbose=True))
