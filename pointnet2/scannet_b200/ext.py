"""The nine functions of the reference's compiled module ``pointnet2._ext``
(/root/reference/pointnet2/_ext_src/src/bindings.cpp:6-19), served by
libpn2_b200.so.

Same names, argument order, dtypes and return shapes as the reference bindings
(ball_query.cpp, group_points.cpp, interpolate.cpp, sampling.cpp).  Like them,
each function checks its tensors (contiguous / float32 / int32 / CUDA --
_ext_src/include/utils.h:5-25) and raises RuntimeError on violation, allocates
its outputs on the inputs' device, and launches on the current stream without
synchronising.  Unlike them it also checks shapes, never calls exit(), and has
no "CPU not supported" assert to trip: CPU tensors are rejected up front.
"""
import torch

from . import lib as _lib


def _req(t, name, dtype, ndim):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s tensor" % (
            name, "a float" if dtype == torch.float32 else "an int"))
    if t.dim() != ndim:
        raise RuntimeError("%s must have %d dimensions, got %s" % (name, ndim, tuple(t.shape)))


def _same_device(*ts):
    d = ts[0].device
    for t in ts[1:]:
        if t.device != d:
            raise RuntimeError("tensors must be on the same device")
    return d


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _f32(shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


def _i32(shape, device):
    return torch.empty(shape, dtype=torch.int32, device=device)


def furthest_point_sampling(points, nsamples):
    """(B, N, 3) f32 -> (B, nsamples) i32      [sampling.cpp:66-87]"""
    _req(points, "points", torch.float32, 3)
    B, N, three = points.shape
    if three != 3:
        raise RuntimeError("points must be (B, N, 3)")
    nsamples = int(nsamples)
    dev = points.device
    out = _i32((B, nsamples), dev)
    temp = None
    if N > 8192:  # PN2_FPS_MAX_RESIDENT_N: streaming kernel keeps the minima here
        temp = _f32((B, N), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_furthest_point_sampling(
            B, N, nsamples, points.data_ptr(), temp.data_ptr() if temp is not None else None,
            out.data_ptr(), _stream(dev)), "furthest_point_sampling")
    return out


def gather_points(points, idx):
    """(B, C, N) f32, (B, m) i32 -> (B, C, m)      [sampling.cpp:15-40]"""
    _req(points, "points", torch.float32, 3)
    _req(idx, "idx", torch.int32, 2)
    dev = _same_device(points, idx)
    B, C, N = points.shape
    m = idx.shape[1]
    out = _f32((B, C, m), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_gather_points(
            B, C, N, m, points.data_ptr(), idx.data_ptr(), out.data_ptr(), _stream(dev)),
            "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """(B, C, m) f32, (B, m) i32, n -> (B, C, n)      [sampling.cpp:42-65]"""
    _req(grad_out, "grad_out", torch.float32, 3)
    _req(idx, "idx", torch.int32, 2)
    dev = _same_device(grad_out, idx)
    B, C, m = grad_out.shape
    out = _f32((B, C, int(n)), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_gather_points_grad(
            B, C, int(n), m, grad_out.data_ptr(), idx.data_ptr(), out.data_ptr(), _stream(dev)),
            "gather_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """centres (B, m, 3), points (B, N, 3) -> (B, m, nsample) i32   [ball_query.cpp:8-32]"""
    _req(new_xyz, "new_xyz", torch.float32, 3)
    _req(xyz, "xyz", torch.float32, 3)
    dev = _same_device(new_xyz, xyz)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    if new_xyz.shape[0] != B or new_xyz.shape[2] != 3 or xyz.shape[2] != 3:
        raise RuntimeError("new_xyz must be (B, m, 3) and xyz (B, N, 3)")
    idx = _i32((B, m, int(nsample)), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_query_ball_point(
            B, N, m, float(radius), int(nsample), new_xyz.data_ptr(), xyz.data_ptr(),
            idx.data_ptr(), _stream(dev)), "ball_query")
    return idx


def group_points(points, idx):
    """(B, C, N) f32, (B, npoint, nsample) i32 -> (B, C, npoint, nsample)  [group_points.cpp:12-36]"""
    _req(points, "points", torch.float32, 3)
    _req(idx, "idx", torch.int32, 3)
    dev = _same_device(points, idx)
    B, C, N = points.shape
    _, npoint, nsample = idx.shape
    out = _f32((B, C, npoint, nsample), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_group_points(
            B, C, N, npoint, nsample, points.data_ptr(), idx.data_ptr(), out.data_ptr(),
            _stream(dev)), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """(B, C, npoint, nsample) f32, idx, n -> (B, C, n)      [group_points.cpp:38-62]"""
    _req(grad_out, "grad_out", torch.float32, 4)
    _req(idx, "idx", torch.int32, 3)
    dev = _same_device(grad_out, idx)
    B, C, npoint, nsample = grad_out.shape
    out = _f32((B, C, int(n)), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_group_points_grad(
            B, C, int(n), npoint, nsample, grad_out.data_ptr(), idx.data_ptr(), out.data_ptr(),
            _stream(dev)), "group_points_grad")
    return out


def three_nn(unknowns, knows):
    """(B, n, 3), (B, m, 3) -> [dist2 (B, n, 3) f32 -- SQUARED --, idx (B, n, 3) i32]
    [interpolate.cpp:14-40]"""
    _req(unknowns, "unknowns", torch.float32, 3)
    _req(knows, "knows", torch.float32, 3)
    dev = _same_device(unknowns, knows)
    B, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = _f32((B, n, 3), dev)
    idx = _i32((B, n, 3), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_three_nn(
            B, n, m, unknowns.data_ptr(), knows.data_ptr(), dist2.data_ptr(), idx.data_ptr(),
            _stream(dev)), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """(B, C, m) f32, (B, n, 3) i32, (B, n, 3) f32 -> (B, C, n)   [interpolate.cpp:42-70]"""
    _req(points, "points", torch.float32, 3)
    _req(idx, "idx", torch.int32, 3)
    _req(weight, "weight", torch.float32, 3)
    dev = _same_device(points, idx, weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = _f32((B, C, n), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_three_interpolate(
            B, C, m, n, points.data_ptr(), idx.data_ptr(), weight.data_ptr(), out.data_ptr(),
            _stream(dev)), "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """(B, C, n) f32, idx, weight, m -> (B, C, m)      [interpolate.cpp:71-99]"""
    _req(grad_out, "grad_out", torch.float32, 3)
    _req(idx, "idx", torch.int32, 3)
    _req(weight, "weight", torch.float32, 3)
    dev = _same_device(grad_out, idx, weight)
    B, C, n = grad_out.shape
    out = _f32((B, C, int(m)), dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().pn2_three_interpolate_grad(
            B, C, n, int(m), grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(),
            out.data_ptr(), _stream(dev)), "three_interpolate_grad")
    return out


__all__ = [
    "gather_points", "gather_points_grad", "furthest_point_sampling",
    "three_nn", "three_interpolate", "three_interpolate_grad",
    "ball_query", "group_points", "group_points_grad",
]
