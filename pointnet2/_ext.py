"""Drop-in for the reference's compiled extension module ``pointnet2._ext``.

The reference's op layer does ``import pointnet2._ext as _ext``
(/root/reference/pointnet2/pointnet2_utils.py:25-33).  ``pointnet2/`` is a
namespace package (no ``__init__.py``) in both trees, so with this repository's
root on ``sys.path`` that import resolves here and the reference's unchanged
``pointnet2_utils.py`` / ``pointnet2_modules.py`` / ``pointnet2_semseg.py`` run
on libpn2_b200.so.
"""
from pointnet2.scannet_b200.ext import (  # noqa: F401
    ball_query, furthest_point_sampling, gather_points, gather_points_grad, group_points,
    group_points_grad, three_interpolate, three_interpolate_grad, three_nn)
