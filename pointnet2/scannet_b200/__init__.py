"""pointnet2.scannet_b200 -- the PointNet++ SA/FP op path of ero1311/Pointnet2.ScanNet on B200.

    ext               the nine functions of the reference's compiled `pointnet2._ext`
    pointnet2_utils   autograd op surface (mirror of the reference's pointnet2_utils.py)
    pointnet2_modules PointnetSAModule(MSG) / PointnetFPModule (+ the fused rows path)
    pytorch_utils     SharedMLP / Conv1d / Conv2d / BatchNorm wrappers (same state-dict keys)
    pointnet2_semseg  Pointnet2SSG / Pointnet2MSG / get_model (the caller, for benchmarks)
    fused, train, profiling, synthetic, build, lib

The kernels live in libpn2_b200.so (C ABI: include/pn2_b200.h), built in-tree by
`python -m pointnet2.scannet_b200.build`.  CUDA only: there is no CPU fallback.
"""
from . import ext, fused, pointnet2_modules, pointnet2_semseg, pointnet2_utils, pytorch_utils  # noqa: F401

__all__ = ["ext", "fused", "pointnet2_modules", "pointnet2_semseg", "pointnet2_utils", "pytorch_utils"]
