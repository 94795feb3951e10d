"""Golden vectors of the reference's per-iteration metrics, made by RUNNING THE REFERENCE:

    python tests/golden/make_metrics_golden.py        # needs /root/reference (dev container only)

Imports /root/reference/scripts/eval.py unmodified (compute_acc, compute_miou, filter_points) with stubs
for what is absent here: lib.config (CONF.NUM_CLASSES = 20, config.py:62), lib.dataset (only imported
names), plyfile (imported by lib/pc_util.py, unused on this path), and the alias ``np.float`` that NumPy
1.24 removed (eval.py:91 uses it; the reference pins an older NumPy).

Cases: "random_labels" -- predictions / targets independent per point: the POINT metrics are pinned
(the reference's voxel metrics depend on Python's per-process hash randomisation there, see
oracle/metrics_np.py); "voxel_constant" -- labels a function of the 2 cm voxel: every metric is pinned.
Inputs are regenerated from the seeds by tests (cases()); outputs go to tests/golden/ref_metrics.npz.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
NUM_CLASSES = 20


def cases():
    """name -> (coords (P, 3) float32, preds, targets int64, weights float32, voxel metrics pinned?)"""
    from pointnet2.scannet_b200 import synthetic
    out = {}
    x, y, w = synthetic.make_batch(2, 2, 4096)                  # with-replacement chunks: many exact duplicates
    coords = np.ascontiguousarray(x[..., :3].reshape(-1, 3))
    rng = np.random.default_rng(5)
    targets = y.reshape(-1) % NUM_CLASSES
    dup_of = {}
    for i, c in enumerate(map(bytes, coords)):                  # duplicates of a point share its target (same scene point)
        targets[i] = targets[dup_of.setdefault(c, i)]
    preds = np.where(rng.random(len(targets)) < 0.6, targets, rng.integers(0, NUM_CLASSES, len(targets)))
    out["random_labels"] = (coords, preds.astype(np.int64), targets.astype(np.int64), w.reshape(-1), False)
    # labels constant per voxel: derive both from the voxel index the reference itself computes
    cmin = coords.min(axis=0)
    v = np.ceil((coords - cmin) / 0.02).astype(np.int64)
    h = v[:, 0] * 73856093 ^ v[:, 1] * 19349663 ^ v[:, 2] * 83492791
    t2 = (h % 7) % NUM_CLASSES                                   # only 7 of the 20 classes are seen
    p2 = np.where((h // 7) % 3 == 0, (t2 + 1) % NUM_CLASSES, t2)
    out["voxel_constant"] = (coords, p2.astype(np.int64), t2.astype(np.int64), w.reshape(-1), True)
    return out


def load_reference_eval():
    cfg = types.ModuleType("lib.config")
    cfg.CONF = types.SimpleNamespace(NUM_CLASSES=NUM_CLASSES)
    ds = types.ModuleType("lib.dataset")
    ds.ScannetDatasetWholeScene = ds.collate_wholescene = None
    ply = types.ModuleType("plyfile")
    ply.PlyData = ply.PlyElement = None
    libpkg = types.ModuleType("lib")
    libpkg.__path__ = [os.path.join(REF, "lib")]               # lib.pc_util is the reference's own file
    sys.modules.update({"lib": libpkg, "lib.config": cfg, "lib.dataset": ds, "plyfile": ply})
    if not hasattr(np, "float"):
        np.float = float
    spec = importlib.util.spec_from_file_location("ref_eval", os.path.join(REF, "scripts", "eval.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ev = load_reference_eval()
    out = {}
    for name, (coords, preds, targets, weights, _) in cases().items():
        pa, papc, va, vapc, vcali, mask = ev.compute_acc(coords, preds, targets, weights)
        pm, vm, mmask = ev.compute_miou(coords, preds, targets, weights)
        for k, v in dict(pointacc=pa, pointacc_per_class=papc, voxacc=va, voxacc_per_class=vapc,
                         voxcaliacc=vcali, acc_mask=mask, pointmiou=pm, voxmiou=vm, miou_mask=mmask).items():
            out[name + "/" + k] = np.asarray(v, dtype=np.float64)
        out[name + "/n_filtered"] = np.asarray(len(ev.filter_points(coords, preds, targets, weights)[0]))
    path = os.path.join(HERE, "ref_metrics.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays")


if __name__ == "__main__":
    main()
