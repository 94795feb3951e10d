"""PointNet++ semantic-segmentation networks -- the CALLER of the hot path, mirrored
from /root/reference/pointnet2/pointnet2_semseg.py (``get_model`` :7-23,
``Pointnet2MSG`` :25-103, ``Pointnet2SSG`` :105-182) so that benchmarks and tests
have the reference's architecture on a box where /root/reference does not exist.
The reference file itself also runs unchanged over this package's ops (see
INTEGRATION.md); parameter names and creation order are identical, so state
dicts and seeded initialisations are interchangeable.

    forward(pointcloud (B, N, 3 + input_channels)) -> (B, N, num_classes)
"""
import torch.nn as nn
import torch.nn.functional as F

from . import fused
from . import pointnet2_utils
from . import pytorch_utils as pt_utils
from .pointnet2_modules import PointnetFPModule, PointnetSAModule, PointnetSAModuleMSG

NPOINTS = [1024, 256, 64, 16]

SSG = dict(
    radius=[0.1, 0.2, 0.4, 0.8],
    nsample=[32, 32, 32, 32],
    mlps=[[32, 32, 64], [64, 64, 128], [128, 128, 256], [256, 256, 512]],
    fp_mlps=[[128, 128], [256, 128], [256, 256], [256, 256]],
)
MSG = dict(
    radius=[[0.05, 0.1], [0.1, 0.2], [0.2, 0.4], [0.4, 0.8]],
    nsample=[[16, 32], [16, 32], [16, 32], [16, 32]],
    mlps=[[[16, 16, 32], [32, 32, 64]], [[64, 64, 128], [64, 96, 128]],
          [[128, 196, 256], [128, 196, 256]], [[256, 256, 512], [256, 384, 512]]],
    fp_mlps=[[128, 128], [256, 256], [512, 512], [512, 512]],
)
CLS_FC = [128]
DP_RATIO = 0.5


class _Pointnet2SemSeg(nn.Module):
    """4 SA levels (8192 -> 1024 -> 256 -> 64 -> 16), 4 FP levels back up, 2-layer head."""

    msg = False
    cfg = SSG

    def __init__(self, num_classes, input_channels=3, use_xyz=True, bn=True):
        super().__init__()
        cfg = self.cfg
        self.SA_modules = nn.ModuleList()
        channel_in = input_channels
        skip_channels = [input_channels]
        for k, npoint in enumerate(NPOINTS):
            if self.msg:
                specs = [[channel_in] + list(s) for s in cfg["mlps"][k]]
                channel_out = sum(s[-1] for s in specs)
                self.SA_modules.append(PointnetSAModuleMSG(
                    npoint=npoint, radii=cfg["radius"][k], nsamples=cfg["nsample"][k],
                    mlps=specs, use_xyz=use_xyz, bn=bn))
            else:
                spec = [channel_in] + list(cfg["mlps"][k])
                channel_out = spec[-1]
                self.SA_modules.append(PointnetSAModule(
                    npoint=npoint, radius=cfg["radius"][k], nsample=cfg["nsample"][k],
                    mlp=spec, use_xyz=use_xyz, bn=bn))
            skip_channels.append(channel_out)
            channel_in = channel_out

        self.FP_modules = nn.ModuleList()
        fp = cfg["fp_mlps"]
        for k in range(len(fp)):
            pre = fp[k + 1][-1] if k + 1 < len(fp) else channel_out
            self.FP_modules.append(PointnetFPModule(mlp=[pre + skip_channels[k]] + fp[k], bn=bn))

        layers = []
        pre = fp[0][-1]
        for width in CLS_FC:
            layers.append(pt_utils.Conv1d(pre, width, bn=bn))
            pre = width
        layers.append(pt_utils.Conv1d(pre, num_classes, activation=None, bn=bn))
        layers.insert(1, nn.Dropout(DP_RATIO))
        self.cls_layer = nn.Sequential(*layers)

    @staticmethod
    def _break_up_pc(pc, views=False):
        """``views``: coordinates and features as column VIEWS of the input rows (no copies) -- for the fused
        training path with a geometry plan, whose kernels read them with the input's row stride."""
        if views and pc.size(-1) > 3:
            return pc[..., 0:3], fused.channel_major(pc[..., 3:])
        xyz = pc[..., 0:3].contiguous()
        if pc.size(-1) <= 3:
            return xyz, None
        if fused.LAZY_CHANNEL_MAJOR:
            # every consumer is a fused block reading point-major rows: one copy of the feature columns,
            # handed on as a channel-major VIEW that remembers its rows (no transpose there and back)
            return xyz, fused.channel_major(pc[..., 3:].contiguous())
        features = pc[..., 3:].transpose(1, 2).contiguous()
        return xyz, features

    def forward(self, pointcloud, plan=None):
        """``plan`` (optional): a geometry.GeometryPlan built from this pointcloud's xyz; the
        sampling / grouping / neighbour indices are then taken from it instead of being
        recomputed inline (identical values, see geometry.py)."""
        prev_lazy = fused.LAZY_CHANNEL_MAJOR
        fused.LAZY_CHANNEL_MAJOR = fused.ENABLED and pointcloud.is_cuda and self._all_fused(pointcloud)
        try:
            return self._forward(pointcloud, plan)
        finally:
            fused.LAZY_CHANNEL_MAJOR = prev_lazy

    def _all_fused(self, pointcloud):
        """True when every SA / FP block takes the fused rows path, i.e. no module output is consumed
        as a channel-major tensor (then those tensors need not be materialised)."""
        like = pointcloud
        if not all(fused.mlp_eligible(mlp, like) for sa in self.SA_modules for mlp in sa.mlps):
            return False
        if not all(fused.mlp_eligible(fp.mlp, like) for fp in self.FP_modules):
            return False
        widths = [sum(mlp[-1].conv.out_channels for mlp in sa.mlps) for sa in self.SA_modules]
        return all(w % 4 == 0 for w in widths) and all(
            isinstance(g, pointnet2_utils.QueryAndGroup) and g.use_xyz and not g.normalize_xyz
            and not g.sample_uniformly and not g.ret_grouped_xyz and not g.ret_unique_cnt
            for sa in self.SA_modules for g in sa.groupers)

    def _forward(self, pointcloud, plan=None):
        # (column views instead of copies where every consumer reads them in place: the training path of the
        # fused blocks with the geometry computed ahead -- SALead / FPLead take the input's row stride)
        views = (plan is not None and fused.LAZY_CHANNEL_MAJOR and fused.LEAD and self.training
                 and pointcloud.is_contiguous() and not pointcloud.requires_grad and pointcloud.size(-1) > 3)
        xyz, features = self._break_up_pc(pointcloud, views)
        l_xyz, l_features = [xyz], [features]
        table0 = None
        if (fused.LAZY_CHANNEL_MAJOR and features is not None and pointcloud.is_contiguous()
                and not pointcloud.requires_grad and pointcloud.dtype == xyz.dtype):
            # the level-0 point table [xyz | features | pad] IS the input, widened to a multiple of 4 columns
            table0 = fused.pad_rows(pointcloud, fused.pad4(pointcloud.size(-1)))
        for k, sa in enumerate(self.SA_modules):
            nxt_xyz, nxt_feat = sa(l_xyz[-1], l_features[-1],
                                   tuple(plan.sa[k]) + (plan.sa_inv[k],) if plan is not None else None,
                                   table0 if k == 0 else None)
            l_xyz.append(nxt_xyz)
            l_features.append(nxt_feat)
        nfp = len(self.FP_modules)
        geo = (lambda k: tuple(plan.fp[k]) + (plan.fp_inv[k],)) if plan is not None else (lambda k: None)
        for i in range(-1, -nfp, -1):  # coarse -> fine, all but the finest level
            l_features[i - 1] = self.FP_modules[i](
                l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i], geo(nfp + i))
        fp0 = self.FP_modules[0]
        if self._head_rows_ok(l_features[1]):
            # finest FP level and the classifier head stay in [points, channels] rows: the head's
            # output rows ARE the (B, N, num_classes) result, no transpose anywhere
            lead, drops, used = self._head_lead(fp0, l_features[0], l_features[1])
            rows = fp0.forward_rows(l_xyz[0], l_xyz[1], l_features[0], l_features[1], geo(0), tail=lead,
                                    tail_dropouts=drops, mark=True)
            return self._head_rows(rows, skip=used)
        l_features[0] = fp0(l_xyz[0], l_xyz[1], l_features[0], l_features[1], geo(0))
        return self.cls_layer(l_features[0]).transpose(1, 2).contiguous()

    def _head_rows_ok(self, known_feats):
        units = [m for m in self.cls_layer if not isinstance(m, nn.Dropout)]
        return (fused.mlp_eligible(units, known_feats) and fused.mlp_eligible(self.FP_modules[0].mlp, known_feats)
                and known_feats.size(1) % 4 == 0)

    def _head_lead(self, fp0, skip_feats, known_feats):
        """The classifier (Conv1d+BN+ReLU, Dropout, Conv1d+BN -- ref :152-159) continues the finest FP
        block's chain as far as the kernels allow: units whose BatchNorms share the block's eps /
        momentum, and a Dropout between two units where both have fused backward layers (it is then a
        keep mask inside the second unit's kernels; in eval mode it is the identity).
        -> (units, {index in units: p of the Dropout after that unit}, modules of cls_layer consumed)"""
        ref = fp0.mlp[0].bn.bn
        mods = list(self.cls_layer)
        rows_shape = (known_feats.size(0) * skip_feats.size(2),
                      fused.pad4(known_feats.size(1) + skip_feats.size(1))) if skip_feats is not None else None
        units, drops, used, pending = [], {}, 0, None
        for i, m in enumerate(mods):
            if isinstance(m, nn.Dropout):
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if not units or nxt is None or isinstance(nxt, nn.Dropout) or pending is not None:
                    break
                if self.training and m.p > 0.0:
                    chain = list(fp0.mlp) + units + [nxt]
                    if rows_shape is None or not fused.chain_supports_dropout(chain, rows_shape, len(chain) - 2):
                        break
                    pending = m.p
                else:
                    pending = 0.0
                continue
            bn = m.bn.bn
            if not (bn.eps == ref.eps and bn.momentum == ref.momentum and m.training == fp0.mlp[0].training):
                break
            if pending:
                drops[len(units) - 1] = pending
            pending = None
            units.append(m)
            used = i + 1
        return units, drops, used

    def _head_rows(self, rows, skip=0):
        """cls_layer (Conv1d+BN+ReLU, Dropout, Conv1d+BN -- ref :152-159) on (B, N, C) rows; the first
        ``skip`` units were already applied.  A Dropout between two units is applied inside the second
        unit's kernels where they support it (the dropped activation is then never materialised)."""
        B, N, C = rows.shape
        x = rows.reshape(B * N, C)
        mods = list(self.cls_layer)[skip:]
        i = 0
        while i < len(mods):
            if isinstance(mods[i], nn.Dropout):
                x = F.dropout(x, mods[i].p, self.training)
                i += 1
                continue
            run, drops = [], {}
            while i < len(mods):
                m = mods[i]
                if isinstance(m, nn.Dropout):
                    nxt = mods[i + 1] if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Dropout) else None
                    if not (run and nxt is not None and self.training and x.is_cuda and
                            fused.chain_supports_dropout(run + [nxt], x, len(run) - 1)):
                        break
                    drops[len(run) - 1] = m.p
                else:
                    run.append(m)
                i += 1
            x = fused.run_mlp_rows(run, x, 0, drops)
        return x.view(B, N, -1)


class Pointnet2SSG(_Pointnet2SemSeg):
    msg = False
    cfg = SSG


class Pointnet2MSG(_Pointnet2SemSeg):
    msg = True
    cfg = MSG


def get_model(num_classes, is_msg=True, input_channels=6, use_xyz=True, bn=True):
    cls = Pointnet2MSG if is_msg else Pointnet2SSG
    return cls(num_classes=num_classes, input_channels=input_channels, use_xyz=use_xyz, bn=bn)
