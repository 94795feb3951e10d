"""Device metrics (csrc/metrics.cu behind pointnet2/scannet_b200/metrics.py) against the golden outputs of
the reference's own compute_acc / compute_miou and against the NumPy oracle, through the C ABI.

Bar: the confusion matrices (integer counts) bit-exact against the oracle; the float64 ratios within
1e-12 relative of the reference / the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_metrics_golden as G  # noqa: E402
from oracle import metrics_np as O  # noqa: E402
from test_metrics_cpu import POINT_KEYS, VOXEL_KEYS, oracle_metrics  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def device_metrics(coords, preds, targets):
    from pointnet2.scannet_b200 import metrics
    c, p, t = (torch.from_numpy(a).to(DEV) for a in (coords, preds, targets))
    pa, papc, va, vapc, vcali, mask = metrics.compute_acc(c, p, t, None, G.NUM_CLASSES)
    pm, vm, mmask = metrics.compute_miou(c, p, t, None, G.NUM_CLASSES)
    d = dict(pointacc=pa, pointacc_per_class=papc, voxacc=va, voxacc_per_class=vapc, voxcaliacc=vcali,
             acc_mask=mask, pointmiou=pm, voxmiou=vm, miou_mask=mmask)
    return {k: v.cpu().numpy() for k, v in d.items()}


@pytest.mark.parametrize("name", ["random_labels", "voxel_constant"])
def test_metrics_equal_reference_and_oracle(name):
    golden = np.load(os.path.join(HERE, "golden", "ref_metrics.npz"))
    coords, preds, targets, weights, voxel_pinned = G.cases()[name]
    got = device_metrics(coords, preds, targets)
    for k in POINT_KEYS + (VOXEL_KEYS if voxel_pinned else ()):
        assert np.allclose(got[k], golden[name + "/" + k], rtol=1e-12, atol=0), k
    want = oracle_metrics(coords, preds, targets, weights)      # same representative rule: everything agrees
    for k in POINT_KEYS + VOXEL_KEYS:
        assert np.allclose(got[k], want[k], rtol=1e-12, atol=0), k


def test_confusion_matrices_full_batch():
    """BASELINE-sized input (32 x 8192 points of with-replacement chunks): integer counts against the oracle;
    size-independent properties: the point matrix counts every distinct coordinate once, the voxel matrix
    every occupied voxel once, and a second call on a shuffled copy gives the same point matrix."""
    from pointnet2.scannet_b200 import metrics, synthetic
    x, y, _ = synthetic.make_batch(2, 32, 8192)
    coords = np.ascontiguousarray(x[..., :3].reshape(-1, 3))
    rng = np.random.default_rng(0)
    targets = (y.reshape(-1) % G.NUM_CLASSES).astype(np.int64)
    first = {}
    for i, c in enumerate(map(bytes, coords)):
        targets[i] = targets[first.setdefault(c, i)]
    preds = np.where(rng.random(len(targets)) < 0.5, targets, rng.integers(0, G.NUM_CLASSES, len(targets)))
    for i, c in enumerate(map(bytes, coords)):                  # a point's duplicates carry its prediction too
        preds[i] = preds[first[c]]
    c, p, t = (torch.from_numpy(a).to(DEV) for a in (coords, preds, targets))
    point, vox, bad = metrics.confusions(c, p, t, G.NUM_CLASSES)
    ones = np.ones(len(targets), np.float32)
    want_point, want_vox = O.confusions(coords, preds, targets, ones, G.NUM_CLASSES)
    assert int(bad.item()) == 0
    assert np.array_equal(point.cpu().numpy(), want_point)
    assert np.array_equal(vox.cpu().numpy(), want_vox)
    assert int(point.sum().item()) == len(first)
    perm = torch.randperm(len(targets), device=DEV)
    point2, _, _ = metrics.confusions(c[perm], p[perm], t[perm], G.NUM_CLASSES)
    assert torch.equal(point, point2)
    logits = torch.nn.functional.one_hot(p, G.NUM_CLASSES).float()
    ev = metrics.evaluate(c, logits, t, G.NUM_CLASSES)
    pa = O.compute_acc(coords, preds, targets, ones, G.NUM_CLASSES)[0]
    assert abs(float(ev["point_acc"].item()) - pa) < 1e-12


def test_metrics_reject_bad_arguments():
    from pointnet2.scannet_b200 import lib, metrics
    L = lib.load()
    assert L.pn2_first_occurrence_xyz(0, None, None, None, None) == -1
    assert L.pn2_confusion(4, None, None, None, 65, None, None, None) == -1
    with pytest.raises(RuntimeError):
        metrics.compute_acc(torch.zeros(4, 3), torch.zeros(4, dtype=torch.long), torch.zeros(4, dtype=torch.long))
