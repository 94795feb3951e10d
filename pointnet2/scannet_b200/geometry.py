"""The weight-independent part of a PointNet++ forward pass, computed ahead of time.

Everything the SA / FP blocks derive from coordinates alone -- furthest point samples, ball-query
neighbourhoods, three nearest neighbours and their inverse-distance weights
(/root/reference/pointnet2/pointnet2_modules.py:54-63 and :394-397) -- depends on the input cloud
but on no parameter.  ``build_plan`` runs that chain once (same kernels, same results as the
modules would produce inline) so that a trainer can compute it for batch i+1 on a side stream
while batch i is in its MLPs (``train.PipelinedTrainStep``): FPS is a latency-bound kernel that
occupies 32 of the 148 SMs, which is exactly what overlaps well with bandwidth-bound layers.
"""
import torch

from . import ext


class GeometryPlan(object):
    """sa[k] = (new_xyz (B, npoint, 3), [idx (B, npoint, nsample) per scale]);
    fp[k] = (idx (B, n, 3) int32, weight (B, n, 3)) for FP_modules[k]."""

    __slots__ = ("sa", "fp", "_flat")

    def __init__(self, sa, fp):
        self.sa, self.fp = sa, fp
        self._flat = None

    def tensors(self):
        out = []
        for new_xyz, idxs in self.sa:
            out.append(new_xyz)
            out.extend(idxs)
        for idx, w in self.fp:
            out.extend([idx, w])
        return out

    def copy_from(self, other):
        """Device copy of another plan of the same shapes: ONE copy when both plans live in flat buffers
        (``pack``), else one per tensor."""
        a, b = getattr(self, "_flat", None), getattr(other, "_flat", None)
        if a is not None and b is not None and a.numel() == b.numel():
            a.copy_(b, non_blocking=True)
            return
        for dst, src in zip(self.tensors(), other.tensors()):
            dst.copy_(src, non_blocking=True)

    def pack(self):
        """Re-home every tensor of the plan as a view into one flat byte buffer (values kept), so that a
        trainer can hand a whole plan over with a single device copy.  -> self"""
        ts = self.tensors()
        offs, total = [], 0
        for t in ts:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        flat = torch.empty(total, dtype=torch.uint8, device=ts[0].device)
        views = []
        for t, o in zip(ts, offs):
            v = flat[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
            v.copy_(t)
            views.append(v)
        it = iter(views)
        self.sa = [(next(it), [next(it) for _ in idxs]) for _, idxs in self.sa]
        self.fp = [(next(it), next(it)) for _ in self.fp]
        self._flat = flat
        return self


@torch.no_grad()
def build_plan(model, xyz):
    """xyz (B, N, 3) -> GeometryPlan for ``model`` (a pointnet2_semseg network of this package)."""
    levels = [xyz]
    sa = []
    for m in model.SA_modules:
        cur = levels[-1]
        picked = ext.furthest_point_sampling(cur, m.npoint)
        new_xyz = ext.gather_points(cur.transpose(1, 2).contiguous(), picked).transpose(1, 2).contiguous()
        idxs = [ext.ball_query(new_xyz, cur, g.radius, g.nsample) for g in m.groupers]
        sa.append((new_xyz, idxs))
        levels.append(new_xyz)
    fp = []
    for k in range(len(model.FP_modules)):
        dist2, idx = ext.three_nn(levels[k], levels[k + 1])
        recip = 1.0 / (torch.sqrt(dist2) + 1e-8)             # pointnet2_modules.py:395-397
        fp.append((idx, recip / torch.sum(recip, dim=2, keepdim=True)))
    return GeometryPlan(sa, fp)
