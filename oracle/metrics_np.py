"""NumPy oracle of the reference's per-iteration metrics -- TEST INFRASTRUCTURE ONLY.

Restates scripts/eval.py:39-123 (filter_points, compute_acc, compute_miou) and lib/pc_util.py:39-51
(point_cloud_label_to_surface_voxel_label_fast) of /root/reference, with one stated difference:

  The reference orders the de-duplicated points by Python's PER-PROCESS-RANDOMISED string hash
  (eval.py:42-43: hash(str(x) + str(y) + str(z)), np.unique sorts by hash value), and a voxel takes the
  labels of the first point of that order (pc_util.py:45-50).  Its voxel metrics therefore change from
  process to process whenever a voxel holds points with different labels.  Here the de-duplicated
  points keep their ORIGINAL order (lowest index first) -- one of the orders the reference may produce.
  The point metrics do not depend on the order; nor do the voxel metrics when the labels are constant
  per voxel, which is how tests/golden/make_metrics_golden.py pins them against the reference itself.

Arithmetic follows the reference: float32 voxel keys (float32 coordinates, weak Python scalar `res`),
integer counts, float64 ratios with the reference's 1e-8 guards.  Only tests/ may import this module.
"""
import numpy as np


def filter_points(coords, preds, targets, weights):
    """eval.py:39-46: first occurrence of every distinct coordinate triple (exact float equality)."""
    keys = np.ascontiguousarray(coords, dtype=np.float32).view(np.uint32).reshape(-1, 3)
    _, first = np.unique(keys, axis=0, return_index=True)
    ids = np.sort(first)                                    # original order (see the module docstring)
    return coords[ids], preds[ids], targets[ids], weights[ids]


def surface_voxel_labels(point_cloud, label, res):
    """lib/pc_util.py:39-51."""
    coordmax = np.max(point_cloud, axis=0)
    coordmin = np.min(point_cloud, axis=0)
    nvox = np.ceil((coordmax - coordmin) / res)
    vidx = np.ceil((point_cloud - coordmin) / res)
    vidx = vidx[:, 0] + vidx[:, 1] * nvox[0] + vidx[:, 2] * nvox[0] * nvox[1]
    uvidx, vpidx = np.unique(vidx, return_index=True)
    return uvidx, label[vpidx, :], nvox


def confusions(coords, preds, targets, weights, num_classes):
    """The two target x prediction count matrices every metric below is a function of."""
    c, p, t, _ = filter_points(coords, preds, targets, weights)
    point = np.zeros((num_classes, num_classes), np.int64)
    np.add.at(point, (t, p), 1)
    _, uv, _ = surface_voxel_labels(c, np.stack([t, p], axis=1), res=0.02)
    vox = np.zeros((num_classes, num_classes), np.int64)
    np.add.at(vox, (uv[:, 0], uv[:, 1]), 1)
    return point, vox


def compute_acc(coords, preds, targets, weights, num_classes):
    """eval.py:48-105 -> pointacc, pointacc_per_class, voxacc, voxacc_per_class, voxcaliacc, mask."""
    point, vox = confusions(coords, preds, targets, weights, num_classes)
    seen = point.sum(1) > 0
    mask = seen.astype(np.float64)
    pointacc = np.trace(point) / float(point.sum())
    voxacc = np.trace(vox) / float(vox.sum())
    labelweights_vox = vox.sum(1).astype(np.float32) / np.sum(vox.sum(1).astype(np.float32))
    correct_vox, seen_vox = np.diag(vox) * seen, vox.sum(1) * seen
    voxcaliacc = np.average(correct_vox / (seen_vox.astype(np.float64) + 1e-8), weights=labelweights_vox)
    pointacc_per_class = np.where(seen, np.diag(point) / (point.sum(1) + 1e-8), 0.0)
    voxacc_per_class = np.where(seen, np.diag(vox) / (vox.sum(1) + 1e-8), 0.0)
    return pointacc, pointacc_per_class, voxacc, voxacc_per_class, voxcaliacc, mask


def compute_miou(coords, preds, targets, weights, num_classes):
    """eval.py:107-129 -> pointmiou, voxmiou, mask (IoU per class: intersection / (union + 1e-8))."""
    point, vox = confusions(coords, preds, targets, weights, num_classes)
    seen = point.sum(1) > 0
    out = []
    for m in (point, vox):
        inter = np.diag(m).astype(np.float64)
        union = m.sum(0) + m.sum(1) - np.diag(m)
        out.append(np.where(seen, inter / (union + 1e-8), 0.0))
    return out[0], out[1], seen.astype(np.float64)
