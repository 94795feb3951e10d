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
    fp[k] = (idx (B, n, 3) int32, weight (B, n, 3)) for FP_modules[k];
    sa_inv[k] = [inverse map (offsets (B, N + 1), items (B, npoint * nsample)) of every scale's idx, or None];
    fp_inv[k] = inverse map (offsets (B, m + 1), items (B, 3 n)) of fp[k]'s idx, or None -- what the
    adjoints of the two row gathers read instead of scattering (csrc/inverse_map.cu)."""

    __slots__ = ("sa", "fp", "sa_inv", "fp_inv", "_flat")

    def __init__(self, sa, fp, sa_inv=None, fp_inv=None):
        self.sa, self.fp = sa, fp
        self.sa_inv = sa_inv if sa_inv is not None else [[None] * len(idxs) for _, idxs in sa]
        self.fp_inv = fp_inv if fp_inv is not None else [None] * len(fp)
        self._flat = None

    def tensors(self):
        out = []
        for new_xyz, idxs in self.sa:
            out.append(new_xyz)
            out.extend(idxs)
        for idx, w in self.fp:
            out.extend([idx, w])
        for invs in self.sa_inv:
            for inv in invs:
                if inv is not None:
                    out.extend(inv)
        for inv in self.fp_inv:
            if inv is not None:
                out.extend(inv)
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
        self.sa_inv = [[(next(it), next(it)) if inv is not None else None for inv in invs] for invs in self.sa_inv]
        self.fp_inv = [(next(it), next(it)) if inv is not None else None for inv in self.fp_inv]
        self._flat = flat
        return self


def inverse_map(keys, n_src):
    """keys (B, ...) int32 with values in [0, n_src) -> (offsets (B, n_src + 1), items (B, n_items)) int32:
    per cloud and key the ascending list of flat positions holding it; None when the sizes are not
    supported by the shared-memory builder (the adjoints then scatter with atomics)."""
    from . import lib
    L = lib.load()
    B = keys.shape[0]
    keys = keys.reshape(B, -1)
    n_items = keys.shape[1]
    if not keys.is_cuda or not L.pn2_inverse_map_supported(n_src, n_items):
        return None
    keys = keys.contiguous()
    offsets = torch.empty(B, n_src + 1, dtype=torch.int32, device=keys.device)
    items = torch.empty(B, n_items, dtype=torch.int32, device=keys.device)
    with torch.cuda.device(keys.device):
        lib.check(L.pn2_inverse_map(B, n_src, n_items, keys.data_ptr(), offsets.data_ptr(), items.data_ptr(),
                                    torch.cuda.current_stream(keys.device).cuda_stream), "inverse_map")
    return offsets, items


@torch.no_grad()
def build_plan(model, xyz, inverse=None):
    """xyz (B, N, 3) -> GeometryPlan for ``model`` (a pointnet2_semseg network of this package);
    ``inverse``: also the inverse maps the gather adjoints read (default: fused.INVERSE_ADJOINT)."""
    if inverse is None:
        from . import fused
        inverse = fused.INVERSE_ADJOINT
    levels = [xyz]
    sa, sa_inv = [], []
    for k, m in enumerate(model.SA_modules):
        cur = levels[-1]
        picked = ext.furthest_point_sampling(cur, m.npoint)
        new_xyz = ext.gather_points(cur.transpose(1, 2).contiguous(), picked).transpose(1, 2).contiguous()
        idxs = [ext.ball_query(new_xyz, cur, g.radius, g.nsample) for g in m.groupers]
        sa.append((new_xyz, idxs))
        # level 0 groups the network's input, which takes no gradient: no adjoint, no inverse map
        sa_inv.append([inverse_map(i, cur.shape[1]) if (k > 0 and inverse) else None for i in idxs])
        levels.append(new_xyz)
    fp, fp_inv = [], []
    for k in range(len(model.FP_modules)):
        dist2, idx = ext.three_nn(levels[k], levels[k + 1])
        recip = 1.0 / (torch.sqrt(dist2) + 1e-8)             # pointnet2_modules.py:395-397
        fp.append((idx, recip / torch.sum(recip, dim=2, keepdim=True)))
        fp_inv.append(inverse_map(idx, levels[k + 1].shape[1]) if inverse else None)
    return GeometryPlan(sa, fp, sa_inv, fp_inv)
