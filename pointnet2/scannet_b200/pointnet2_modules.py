"""Set-abstraction and feature-propagation blocks -- mirror of
/root/reference/pointnet2/pointnet2_modules.py for the classes the semseg models
use: ``PointnetSAModuleMSG`` (ref :78-124), ``PointnetSAModule`` (:127-161) and
``PointnetFPModule`` (:356-416).  Constructor keywords, child-module names
(``groupers``, ``mlps``, ``mlp``) and tensor layouts are the reference's; the
VoteNet leftovers (``*Votes``, ``PointnetLFPModuleMSG``) are out of scope.

    SA:  xyz (B, N, 3), features (B, C, N)
         -> new_xyz (B, npoint, 3), new_features (B, sum_k mlps[k][-1], npoint)
    FP:  unknown (B, n, 3), known (B, m, 3), unknow_feats (B, C1, n), known_feats (B, C2, m)
         -> (B, mlp[-1], n)
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused
from . import pointnet2_utils
from . import pytorch_utils as pt_utils


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz, features=None, geometry=None, table=None):
        """``geometry`` (optional) = (new_xyz, [idx per scale][, [inverse map of idx per scale]]) computed ahead
        by geometry.build_plan; ``table`` (optional): the [xyz | features | pad] point rows (fused path only)."""
        if geometry is not None and self._fused_ok(xyz, features):
            return geometry[0], self._forward_rows(xyz, geometry[0], features, geometry[1],
                                                   geometry[2] if len(geometry) > 2 else None, table)
        new_xyz = None
        if self.npoint is not None:
            # FPS on (B, N, 3); the centres are gathered from the channel-major copy
            picked = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
            new_xyz = pointnet2_utils.gather_operation(
                xyz.transpose(1, 2).contiguous(), picked).transpose(1, 2).contiguous()

        if self._fused_ok(xyz, features):
            return new_xyz, self._forward_rows(xyz, new_xyz, features, table=table)

        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            grouped = grouper(xyz, new_xyz, features)        # (B, C, npoint, nsample)
            grouped = mlp(grouped)                           # (B, mlp[-1], npoint, nsample)
            grouped = F.max_pool2d(grouped, kernel_size=[1, grouped.size(3)])
            pooled.append(grouped.squeeze(-1))               # (B, mlp[-1], npoint)
        return new_xyz, torch.cat(pooled, dim=1)

    # -- fused path: same arithmetic on [positions, channels] rows (fused.py) -----------
    def _fused_ok(self, xyz, features):
        if self.npoint is None or (features is not None and features.dtype != torch.float32):
            return False
        for grouper, mlp in zip(self.groupers, self.mlps):
            if not isinstance(grouper, pointnet2_utils.QueryAndGroup):
                return False
            if (not grouper.use_xyz or grouper.normalize_xyz or grouper.sample_uniformly
                    or grouper.ret_grouped_xyz or grouper.ret_unique_cnt):
                return False
            if not fused.mlp_eligible(mlp, xyz):
                return False
        return True

    def _forward_rows(self, xyz, new_xyz, features, idxs=None, invs=None, table=None):
        """``table`` (optional): the (B, N, pad4(3 + C)) [xyz | features | 0] rows, when the caller has them."""
        B, N, _ = xyz.shape
        feat_rows = fused.rows_of(features) if features is not None else None

        def the_table():   # built on first use: the training path of levels whose C is a multiple of 4 never needs it
            nonlocal table
            if table is None:
                cols = [xyz] if feat_rows is None else [xyz, feat_rows]
                width = sum(c.size(2) for c in cols)
                if fused.pad4(width) != width:
                    cols.append(xyz.new_zeros(B, N, fused.pad4(width) - width))
                table = torch.cat(cols, dim=2)                   # (B, N, 3 + C [+ pad]) per-point rows
            return table
        pooled = []
        for i, (grouper, mlp) in enumerate(zip(self.groupers, self.mlps)):
            idx = idxs[i] if idxs is not None else pointnet2_utils.ball_query(
                grouper.radius, grouper.nsample, xyz, new_xyz)
            out = None
            if feat_rows is not None and fused.lead_ok(mlp):
                # training: the first conv runs over the N points, the gather writes its output (fused.SALead)
                if feat_rows.size(2) % 4 == 0 and table is None:
                    out = fused.run_mlp_rows(mlp, feat_rows, grouper.nsample, lead=fused.SALead(xyz, new_xyz, idx, 0))
                else:
                    out = fused.run_mlp_rows(mlp, the_table(), grouper.nsample, lead=fused.SALead(xyz, new_xyz, idx, 3))
            if out is None:
                out = fused.sa_chain_eval(mlp, the_table(), new_xyz, idx)    # inference: the whole block in one kernel
            if out is None:
                inv = invs[i] if invs is not None and invs[i] is not None else (None, None)
                rows = fused.SAGroupRows.apply(the_table(), new_xyz, idx, inv[0], inv[1])   # (B*npoint*nsample, ct)
                out = fused.run_mlp_rows(mlp, rows, grouper.nsample)      # (B*npoint, mlp[-1])
            pooled.append(out.view(B, self.npoint, -1))
        rows = torch.cat(pooled, dim=2) if len(pooled) > 1 else pooled[0]   # (B, npoint, sum C)
        return fused.channel_major(rows)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale grouping: one (radius, nsample, mlp) triple per scale, same centres.

    As in the reference (ref :121-122) ``mlps[i][0]`` is incremented IN PLACE by 3
    when ``use_xyz``.
    """

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int],
                 mlps: List[List[int]], bn: bool = True, use_xyz: bool = True,
                 sample_uniformly: bool = False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            if npoint is not None:
                self.groupers.append(pointnet2_utils.QueryAndGroup(
                    radius, nsample, use_xyz=use_xyz, sample_uniformly=sample_uniformly))
            else:
                self.groupers.append(pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3
            self.mlps.append(pt_utils.SharedMLP(spec, bn=bn))


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None,
                 nsample: int = None, bn: bool = True, use_xyz: bool = True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample],
                         bn=bn, use_xyz=use_xyz)


class PointnetFPModule(nn.Module):
    """Feature propagation: inverse-distance interpolation from the 3 nearest known
    points, concatenation with the skip features, shared MLP."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats, geometry=None):
        """``geometry`` (optional) = (idx, weight) of the three nearest known points, computed ahead."""
        if known is not None and fused.mlp_eligible(self.mlp, known_feats) \
                and known_feats.size(1) % 4 == 0:
            return fused.channel_major(self.forward_rows(unknown, known, unknow_feats, known_feats, geometry))

        if known is None:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        else:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            recip = 1.0 / (dist + 1e-8)                      # ref :395-397
            weight = recip / torch.sum(recip, dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)

        feats = interpolated if unknow_feats is None else torch.cat(
            [interpolated, unknow_feats], dim=1)             # (B, C2 + C1, n)
        return self.mlp(feats.unsqueeze(-1)).squeeze(-1)

    def forward_rows(self, unknown, known, unknow_feats, known_feats, geometry=None, tail=(), tail_dropouts=None,
                     mark=False):
        """Fused path, point-major result (B, n, mlp[-1]): interpolation + concat written straight
        into [positions, channels] rows, then the shared MLP on rows (fused.py).  ``tail``: further
        conv + BN (+ ReLU) units applied to the result in the same chain (the classifier's first unit:
        this block's last activation is then never written out or re-read); ``tail_dropouts``:
        {index into tail: p} -- nn.Dropout(p) after that tail unit, applied inside the next unit."""
        B, n, _ = unknown.shape
        inv = (None, None)
        if geometry is not None:
            idx, weight = geometry[0], geometry[1]
            if len(geometry) > 2 and geometry[2] is not None:
                inv = geometry[2]
        else:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            recip = 1.0 / (dist + 1e-8)                      # ref :395-397
            weight = recip / torch.sum(recip, dim=2, keepdim=True)
        skip = fused.rows_of(unknow_feats) if unknow_feats is not None else None
        drops = {len(self.mlp) + i: p for i, p in (tail_dropouts or {}).items()}
        units = list(self.mlp) + list(tail)
        if (fused.lead_ok(units, 64) and (skip is None or (skip.size(2) <= fused.fp_lead_max_skip()
                                                       and not skip.requires_grad))):
            # training, few skip channels (the finest level: the input features): the first conv runs over the m
            # known points and the interpolation writes its output
            return fused.run_mlp_rows(units, fused.rows_of(known_feats), 0, drops,
                                      lead=fused.FPLead(skip, idx, weight, mark)).view(B, n, -1)
        if skip is not None:
            skip = skip.contiguous()       # (a column view of the input rows when the caller passed one)
        rows = fused.FPInterpRows.apply(fused.rows_of(known_feats), skip, idx, weight, mark, inv[0], inv[1])
        return fused.run_mlp_rows(units, rows, 0, drops).view(B, n, -1)
