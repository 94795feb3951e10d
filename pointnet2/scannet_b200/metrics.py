"""Per-iteration evaluation metrics on the device -- mirror of ``compute_acc`` / ``compute_miou``
(/root/reference/scripts/eval.py:48-129, called for every training and validation iteration by
lib/solver.py:327-347), served by csrc/metrics.cu.

The reference moves coordinates, predictions, targets and weights to the host each iteration and runs a
Python loop over all B * N points (``hash(str(x) + str(y) + str(z))``, eval.py:42) before NumPy takes
over; at 32 x 8192 points that costs seconds per iteration next to a training step of milliseconds.
Here the same quantities come from two confusion matrices built on the device (exact-duplicate removal
and 2 cm voxel representatives by lock-free hash tables) and nothing is synchronised: every return value
is a CUDA tensor (float64, the reference's arithmetic).  Same function names, argument order and return
tuples as the reference; ``num_classes`` replaces the global ``CONF.NUM_CLASSES``.

One stated difference: the reference visits the de-duplicated points in the order of Python's
per-process-randomised string hash, and a voxel takes the labels of the first point in that order, so
its voxel metrics vary from process to process when a voxel holds differently labelled points.  Here the
representative is the point with the lowest index (oracle/metrics_np.py pins both against the reference).

There is no CPU path.
"""
import torch

from . import lib as _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def confusions(coords, preds, targets, num_classes, res=0.02):
    """coords (P, 3) float32, preds / targets (P,) int64 on one CUDA device -> (point, voxel) confusion
    matrices [num_classes, num_classes] int32 (rows: target, columns: prediction)."""
    if not (coords.is_cuda and preds.is_cuda and targets.is_cuda):
        raise RuntimeError("metrics need CUDA tensors (CPU not supported)")
    coords = coords.reshape(-1, 3).contiguous().float()
    preds = preds.reshape(-1).contiguous().long()
    targets = targets.reshape(-1).contiguous().long()
    P, dev = coords.shape[0], coords.device
    if preds.numel() != P or targets.numel() != P:
        raise RuntimeError("coords, preds and targets must describe the same points")
    L = _lib.load()
    i32 = dict(dtype=torch.int32, device=dev)
    table = torch.empty(L.pn2_metrics_table_ints(P), **i32)
    rep, vrep, vkey, bounds = torch.empty(P, **i32), torch.empty(P, **i32), torch.empty(P, **i32), torch.empty(6, **i32)
    conf = torch.zeros(2, num_classes, num_classes, **i32)
    bad = torch.zeros(1, **i32)
    s = _stream(dev)
    with torch.cuda.device(dev):
        _lib.check(L.pn2_first_occurrence_xyz(P, coords.data_ptr(), table.data_ptr(), rep.data_ptr(), s),
                   "first_occurrence_xyz")
        _lib.check(L.pn2_confusion(P, rep.data_ptr(), preds.data_ptr(), targets.data_ptr(), num_classes,
                                   conf[0].data_ptr(), bad.data_ptr(), s), "confusion")
        _lib.check(L.pn2_first_occurrence_voxel(P, coords.data_ptr(), rep.data_ptr(), float(res), bounds.data_ptr(),
                                                vkey.data_ptr(), table.data_ptr(), vrep.data_ptr(), s),
                   "first_occurrence_voxel")
        _lib.check(L.pn2_confusion(P, vrep.data_ptr(), preds.data_ptr(), targets.data_ptr(), num_classes,
                                   conf[1].data_ptr(), bad.data_ptr(), s), "confusion")
    return conf[0], conf[1], bad


def _acc(point, vox):
    point, vox = point.double(), vox.double()
    seen = point.sum(1) > 0
    mask = seen.double()
    pointacc = point.diagonal().sum() / point.sum()
    voxacc = vox.diagonal().sum() / vox.sum()
    lw = vox.sum(1).float() / vox.sum(1).float().sum()                    # float32, eval.py:88-89
    ratio = (vox.diagonal() * mask) / (vox.sum(1) * mask + 1e-8)
    voxcaliacc = (ratio * lw.double()).sum() / lw.double().sum()          # np.average(..., weights=)
    zero = torch.zeros_like(mask)
    pointacc_per_class = torch.where(seen, point.diagonal() / (point.sum(1) + 1e-8), zero)
    voxacc_per_class = torch.where(seen, vox.diagonal() / (vox.sum(1) + 1e-8), zero)
    return pointacc, pointacc_per_class, voxacc, voxacc_per_class, voxcaliacc, mask


def _miou(point, vox):
    seen = point.sum(1) > 0
    out = []
    for m in (point.double(), vox.double()):
        union = m.sum(0) + m.sum(1) - m.diagonal()
        out.append(torch.where(seen, m.diagonal() / (union + 1e-8), torch.zeros_like(union)))
    return out[0], out[1], seen.double()


def compute_acc(coords, preds, targets, weights=None, num_classes=20):
    """eval.py:48-105 -> (pointacc, pointacc_per_class, voxacc, voxacc_per_class, voxcaliacc, mask)."""
    point, vox, _ = confusions(coords, preds, targets, num_classes)
    return _acc(point, vox)


def compute_miou(coords, preds, targets, weights=None, num_classes=20):
    """eval.py:107-129 -> (pointmiou, voxmiou, mask)."""
    point, vox, _ = confusions(coords, preds, targets, num_classes)
    return _miou(point, vox)


def evaluate(coords, logits, targets, num_classes=20):
    """Solver._eval (lib/solver.py:327-347) in one pass: arg-max of the logits, both confusion matrices
    once, the six running-log scalars as CUDA tensors (read them every N iterations, or never)."""
    preds = logits.reshape(-1, logits.shape[-1]).argmax(dim=1)
    point, vox, _ = confusions(coords, preds, targets, num_classes)
    pa, papc, va, vapc, _, mask = _acc(point, vox)
    pm, vm, _ = _miou(point, vox)
    n = mask.sum()
    return {"point_acc": pa, "point_acc_per_class": (papc * mask).sum() / n, "voxel_acc": va,
            "voxel_acc_per_class": (vapc * mask).sum() / n, "point_miou": (pm * mask).sum() / n,
            "voxel_miou": (vm * mask).sum() / n}
