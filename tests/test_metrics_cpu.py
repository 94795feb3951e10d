"""oracle/metrics_np.py (NumPy restatement of scripts/eval.py's compute_acc / compute_miou) against golden
outputs of the reference's own functions (tests/golden/make_metrics_golden.py, run on /root/reference)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_metrics_golden as G  # noqa: E402  (case generator only)
from oracle import metrics_np as O  # noqa: E402

POINT_KEYS = ("pointacc", "pointacc_per_class", "acc_mask", "pointmiou", "miou_mask")
VOXEL_KEYS = ("voxacc", "voxacc_per_class", "voxcaliacc", "voxmiou")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "ref_metrics.npz"))


def oracle_metrics(coords, preds, targets, weights):
    pa, papc, va, vapc, vcali, mask = O.compute_acc(coords, preds, targets, weights, G.NUM_CLASSES)
    pm, vm, mmask = O.compute_miou(coords, preds, targets, weights, G.NUM_CLASSES)
    return dict(pointacc=pa, pointacc_per_class=papc, voxacc=va, voxacc_per_class=vapc, voxcaliacc=vcali,
                acc_mask=mask, pointmiou=pm, voxmiou=vm, miou_mask=mmask)


@pytest.mark.parametrize("name", ["random_labels", "voxel_constant"])
def test_oracle_equals_reference(golden, name):
    coords, preds, targets, weights, voxel_pinned = G.cases()[name]
    got = oracle_metrics(coords, preds, targets, weights)
    assert len(O.filter_points(coords, preds, targets, weights)[0]) == int(golden[name + "/n_filtered"])
    for k in POINT_KEYS + (VOXEL_KEYS if voxel_pinned else ()):
        assert np.allclose(got[k], golden[name + "/" + k], rtol=1e-12, atol=0), k
    if not voxel_pinned:   # order-dependent in the reference (randomised hash): same population, close values
        assert abs(got["voxacc"] - float(golden[name + "/voxacc"])) < 0.05
