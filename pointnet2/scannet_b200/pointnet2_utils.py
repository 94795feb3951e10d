"""Autograd op surface of the PointNet++ SA/FP path -- host-side mirror of
/root/reference/pointnet2/pointnet2_utils.py over libpn2_b200.so.

Same public names, positional signatures, dtypes and differentiability as the
reference module, so code written against it (the reference's
``pointnet2_modules.py`` / ``pointnet2_semseg.py``, or this package's mirrors)
runs unchanged:

    furthest_point_sample(xyz, npoint)              ref :51-80   non-differentiable
    gather_operation(features, idx)                 ref :83-117
    three_nn(unknown, known) -> (dist, idx)         ref :120-149 no gradient
    three_interpolate(features, idx, weight)        ref :152-206 grad to features only
    grouping_operation(features, idx)               ref :209-257
    ball_query(radius, nsample, xyz, new_xyz)       ref :260-291 non-differentiable
    QueryAndGroup, GroupAll, RandomDropout          ref :294-425, :40-48

Like the reference (ref :25-33) the kernels are reached through a module-level
``_ext`` object with the nine functions of bindings.cpp:6-19; here it is
``pointnet2.scannet_b200.ext``.  It is CUDA-only: there is no CPU path.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ext as _ext  # swap with `pointnet2_utils._ext = other` (same seam as the reference)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        """xyz (B, N, 3) -> (B, npoint) int32 indices of the sampled set."""
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B, C, N), idx (B, npoint) -> (B, C, npoint)."""
        ctx.for_backwards = (idx, features.size(1), features.size(2))
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, _, N = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, N), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        """unknown (B, n, 3), known (B, m, 3) -> dist (B, n, 3) L2 (not squared), idx (B, n, 3)."""
        dist2, idx = _ext.three_nn(unknown, known)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        """features (B, c, m), idx (B, n, 3), weight (B, n, 3) -> (B, c, n)."""
        ctx.three_interpolate_for_backward = (idx, weight, features.size(2))
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B, C, N), idx (B, npoint, nsample) -> (B, C, npoint, nsample)."""
        ctx.for_backwards = (idx, features.size(2))
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, N), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        """Python order (radius, nsample, xyz, new_xyz); the extension's is
        (new_xyz, xyz, radius, nsample) -- ref :262,282."""
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Ball query around ``new_xyz`` then grouping of centred xyz and features.

    forward(xyz (B, N, 3), new_xyz (B, npoint, 3), features (B, C, N) or None)
      -> (B, 3 + C, npoint, nsample), xyz channels first (ref :317-376).
    """

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False,
                 normalize_xyz=False, sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if ret_unique_cnt:
            assert sample_uniformly

    def _resample_uniformly(self, idx):
        # ref :336-345 -- host loop with device syncs; off by default, unused by semseg
        unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
        for b in range(idx.shape[0]):
            for r in range(idx.shape[1]):
                uniq = torch.unique(idx[b, r, :])
                k = uniq.shape[0]
                unique_cnt[b, r] = k
                fill = torch.randint(0, k, (self.nsample - k,), dtype=torch.long)
                idx[b, r, :] = torch.cat((uniq, uniq[fill]))
        return unique_cnt

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = self._resample_uniformly(idx) if self.sample_uniformly else None

        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)  # centre on the query point
        if self.normalize_xyz:
            grouped_xyz /= self.radius

        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        else:
            grouped = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped

        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """``npoint=None`` grouper: one group holding every point (views only).

    The reference accepts ``ret_grouped_xyz`` but never stores it (ref :387-390),
    so its forward raises AttributeError at ref :422; here the flag is kept.
    """

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            new_features = grouped_xyz
        else:
            grouped = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features


class RandomDropout(nn.Module):
    """Kept for surface completeness.  The reference's forward calls
    ``pt_utils.feature_dropout_no_scaling`` which does not exist (ref :40-48); here
    it is the no-rescale feature dropout that name describes."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p = p
        self.inplace = inplace

    def forward(self, X):
        theta = torch.empty(1).uniform_(0, self.p).item()
        if not self.training or theta == 0.0:
            return X
        keep = (torch.rand(X.shape[:2], device=X.device) >= theta).to(X.dtype)
        keep = keep.view(*X.shape[:2], *([1] * (X.dim() - 2)))
        return X.mul_(keep) if self.inplace else X * keep
