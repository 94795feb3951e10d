"""CPU oracle for the PointNet++ SA/FP op path -- TEST INFRASTRUCTURE ONLY.

``oracle/pn2_oracle.c`` restates the nine reference kernels
(/root/reference/pointnet2/_ext_src/src/*_gpu.cu) in scalar C with FMA
contraction disabled.  This module loads it through ctypes and exposes

* ``ext_cpu`` -- an object with the nine functions of the reference's compiled
  module ``pointnet2._ext`` (_ext_src/src/bindings.cpp:6-19), same names,
  argument order and return shapes, operating on CPU torch tensors.  Assigning
  it to ``pointnet2_utils._ext`` lets the reference's unchanged Python model
  code run on the host (BASELINE.json configs[0]).
* ``block_threads`` / ``fps_tiebreak_key`` -- the launch-geometry helper and the
  closed-form FPS tie-break key the CUDA kernel uses, for cross-checking.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``pointnet2.scannet_b200``) never does; it has no CPU fallback.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpn2_oracle.so")
_lib = None

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)
_i = ctypes.c_int


def build(force=False):
    """Compile oracle/pn2_oracle.c with oracle/Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "pn2_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libpn2_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        sig = {
            "pn2o_furthest_point_sampling": [_i, _i, _i, _F, _F, _I],
            "pn2o_gather_points": [_i, _i, _i, _i, _F, _I, _F],
            "pn2o_gather_points_grad": [_i, _i, _i, _i, _F, _I, _F],
            "pn2o_query_ball_point": [_i, _i, _i, ctypes.c_float, _i, _F, _F, _I],
            "pn2o_group_points": [_i, _i, _i, _i, _i, _F, _I, _F],
            "pn2o_group_points_grad": [_i, _i, _i, _i, _i, _F, _I, _F],
            "pn2o_three_nn": [_i, _i, _i, _F, _F, _F, _I],
            "pn2o_three_interpolate": [_i, _i, _i, _i, _F, _I, _F, _F],
            "pn2o_three_interpolate_grad": [_i, _i, _i, _i, _F, _I, _F, _F],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _i
        L.pn2o_block_threads.argtypes = [_i]
        L.pn2o_block_threads.restype = _i
        L.pn2o_set_threads.argtypes = [_i]
        L.pn2o_set_threads.restype = None
        L.pn2o_get_threads.restype = _i
        _lib = L
    return _lib


def set_threads(n):
    """Clouds of a batch are processed by up to ``n`` host threads."""
    lib().pn2o_set_threads(int(n))


def get_threads():
    return lib().pn2o_get_threads()


def block_threads(work):
    """opt_n_threads of _ext_src/include/cuda_utils.h:15-19."""
    return lib().pn2o_block_threads(int(work))


def fps_tiebreak_key(n):
    """Closed form of the reference FPS arg-max tie-break.

    Among points of equal maximal distance the reference's strided scan plus
    shared-memory tree (sampling_gpu.cu:59-65,90-167) selects the point that
    minimises ``(bitrev_{log2 bs}(k mod bs), k div bs)``.  Returns that key as
    one int64 per point index k (smaller wins).
    """
    bs = block_threads(n)
    bits = bs.bit_length() - 1
    k = np.arange(n, dtype=np.int64)
    t = k % bs
    rev = np.zeros_like(t)
    for i in range(bits):
        rev |= ((t >> i) & 1) << (bits - 1 - i)
    q = (n + bs - 1) // bs
    return rev * q + k // bs


def _fp(t):
    return ctypes.cast(t.data_ptr(), _F)


def _ip(t):
    return ctypes.cast(t.data_ptr(), _I)


def _chk(t, dtype, what):
    if t.is_cuda:
        raise RuntimeError("oracle is CPU-only: %s is a CUDA tensor" % what)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % what)
    if t.dtype != dtype:
        raise RuntimeError("%s must be a %s tensor" % (what, dtype))


class _ExtCPU(object):
    """The nine functions of ``pointnet2._ext`` (bindings.cpp:6-19) on the CPU."""

    @staticmethod
    def furthest_point_sampling(points, nsamples):  # sampling.cpp:66-87
        _chk(points, torch.float32, "points")
        B, N, _ = points.shape
        out = torch.zeros(B, nsamples, dtype=torch.int32)
        lib().pn2o_furthest_point_sampling(B, N, nsamples, _fp(points), None, _ip(out))
        return out

    @staticmethod
    def gather_points(points, idx):  # sampling.cpp:15-40
        _chk(points, torch.float32, "points"), _chk(idx, torch.int32, "idx")
        B, C, N = points.shape
        m = idx.shape[1]
        out = torch.zeros(B, C, m, dtype=torch.float32)
        lib().pn2o_gather_points(B, C, N, m, _fp(points), _ip(idx), _fp(out))
        return out

    @staticmethod
    def gather_points_grad(grad_out, idx, n):  # sampling.cpp:42-65
        _chk(grad_out, torch.float32, "grad_out"), _chk(idx, torch.int32, "idx")
        B, C, m = grad_out.shape
        out = torch.zeros(B, C, n, dtype=torch.float32)
        lib().pn2o_gather_points_grad(B, C, n, m, _fp(grad_out), _ip(idx), _fp(out))
        return out

    @staticmethod
    def ball_query(new_xyz, xyz, radius, nsample):  # ball_query.cpp:8-32
        _chk(new_xyz, torch.float32, "new_xyz"), _chk(xyz, torch.float32, "xyz")
        B, N, _ = xyz.shape
        m = new_xyz.shape[1]
        idx = torch.zeros(B, m, nsample, dtype=torch.int32)
        lib().pn2o_query_ball_point(B, N, m, float(radius), int(nsample),
                                    _fp(new_xyz), _fp(xyz), _ip(idx))
        return idx

    @staticmethod
    def group_points(points, idx):  # group_points.cpp:12-36
        _chk(points, torch.float32, "points"), _chk(idx, torch.int32, "idx")
        B, C, N = points.shape
        _, npoints, nsample = idx.shape
        out = torch.zeros(B, C, npoints, nsample, dtype=torch.float32)
        lib().pn2o_group_points(B, C, N, npoints, nsample, _fp(points), _ip(idx), _fp(out))
        return out

    @staticmethod
    def group_points_grad(grad_out, idx, n):  # group_points.cpp:38-62
        _chk(grad_out, torch.float32, "grad_out"), _chk(idx, torch.int32, "idx")
        B, C, npoints, nsample = grad_out.shape
        out = torch.zeros(B, C, n, dtype=torch.float32)
        lib().pn2o_group_points_grad(B, C, n, npoints, nsample, _fp(grad_out), _ip(idx), _fp(out))
        return out

    @staticmethod
    def three_nn(unknowns, knows):  # interpolate.cpp:14-40 -- returns SQUARED distances
        _chk(unknowns, torch.float32, "unknowns"), _chk(knows, torch.float32, "knows")
        B, n, _ = unknowns.shape
        m = knows.shape[1]
        idx = torch.zeros(B, n, 3, dtype=torch.int32)
        dist2 = torch.zeros(B, n, 3, dtype=torch.float32)
        lib().pn2o_three_nn(B, n, m, _fp(unknowns), _fp(knows), _fp(dist2), _ip(idx))
        return [dist2, idx]

    @staticmethod
    def three_interpolate(points, idx, weight):  # interpolate.cpp:42-70
        _chk(points, torch.float32, "points"), _chk(idx, torch.int32, "idx")
        _chk(weight, torch.float32, "weight")
        B, C, m = points.shape
        n = idx.shape[1]
        out = torch.zeros(B, C, n, dtype=torch.float32)
        lib().pn2o_three_interpolate(B, C, m, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
        return out

    @staticmethod
    def three_interpolate_grad(grad_out, idx, weight, m):  # interpolate.cpp:71-99
        _chk(grad_out, torch.float32, "grad_out"), _chk(idx, torch.int32, "idx")
        _chk(weight, torch.float32, "weight")
        B, C, n = grad_out.shape
        out = torch.zeros(B, C, m, dtype=torch.float32)
        lib().pn2o_three_interpolate_grad(B, C, n, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(out))
        return out


ext_cpu = _ExtCPU()

EXT_FUNCTIONS = (
    "gather_points", "gather_points_grad", "furthest_point_sampling",
    "three_nn", "three_interpolate", "three_interpolate_grad",
    "ball_query", "group_points", "group_points_grad",
)
