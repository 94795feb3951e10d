"""Golden vectors of the reference's chunk generator, made by RUNNING THE REFERENCE:

    python tests/golden/make_chunker_golden.py        # needs /root/reference (dev container only)

Imports /root/reference/lib/dataset.py unmodified.  Its third-party imports that are absent here
(h5py, prefetch_generator, easydict via lib.config) are stubbed in sys.modules: `background()` becomes
the identity decorator, CONF only provides the .npy path template, NUM_CLASSES and the multiview
file name, and h5py.File returns a dict of the synthetic feature matrices.  The scenes are
synthetic (pointnet2/scannet_b200/synthetic.make_prep_scene, seeded) and written as .npy files to a
temporary directory, which is what ScannetDataset._prepare_weights reads.

For every case np.random is seeded, the reference's generate_chunks / __getitem__ /
collate_random (or ScannetDatasetWholeScene.__getitem__) run, and the outputs are stored in
tests/golden/ref_chunker.npz.  tests/test_chunker_cpu.py checks oracle/chunker_np.py against them;
tests/test_chunker_gpu.py checks the CUDA path against both.
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import synthetic  # noqa: E402

REF = "/root/reference"
# (name, scene seeds, scene points, npoints, flags, numpy seed)
TRAIN_CASES = [
    ("train_xyz_rgb_normal", [11, 12, 13, 14], 30000, 2048, dict(use_color=True, use_normal=True), 101),
    ("train_xyz_only", [15, 16], 20000, 1024, dict(), 102),
    ("train_multiview", [17, 18], 20000, 1024, dict(use_multiview=True, use_normal=True), 103),
    ("val_rgb", [19, 20], 20000, 1024, dict(use_color=True), 104),
]
SCENE_CASES = [
    ("scene_rgb_normal", 21, 30000, 1024, dict(use_color=True, use_normal=True), 201),
    ("scene_multiview", 22, 20000, 512, dict(use_normal=True, use_multiview=True), 202),
]
MULTIVIEW_CH = 8  # the reference takes whatever width the h5 rows have (128 in production)


def scenes_for(seeds, n, multiview):
    out = {}
    for s in seeds:
        r = synthetic.make_prep_scene(s, npoints=n, multiview=MULTIVIEW_CH if multiview else 0)
        out["scene%04d_00" % s] = r if multiview else (r, None)
    return out


def load_reference_dataset(tmp, features):
    class _Row(object):
        def __init__(self, a):
            self.a = a

        def __getitem__(self, i):  # `database.get(scene_id)[()]`, lib/dataset.py:46
            return self.a

    class FakeH5(dict):
        def get(self, k):
            return _Row(self[k])

    h5py = types.ModuleType("h5py")
    h5py.File = lambda *a, **k: FakeH5(features)
    pg = types.ModuleType("prefetch_generator")
    pg.background = lambda *a, **k: (lambda f: f)
    conf = types.SimpleNamespace(SCANNETV2_FILE=os.path.join(tmp, "{}.npy"), MULTIVIEW="unused", NUM_CLASSES=21)
    cfg = types.ModuleType("lib.config")
    cfg.CONF = conf
    libpkg = types.ModuleType("lib")
    libpkg.__path__ = []
    sys.modules.update({"h5py": h5py, "prefetch_generator": pg, "lib": libpkg, "lib.config": cfg})
    spec = importlib.util.spec_from_file_location("ref_dataset", os.path.join(REF, "lib", "dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, seeds, n, npoints, flags, seed in TRAIN_CASES:
            sc = scenes_for(seeds, n, flags.get("use_multiview", False))
            for k, (scene, _) in sc.items():
                np.save(os.path.join(tmp, k + ".npy"), scene)
            ds_mod = load_reference_dataset(tmp, {k: f for k, (_, f) in sc.items()})
            phase = "val" if name.startswith("val") else "train"
            ds = ds_mod.ScannetDataset(phase, sorted(sc), num_classes=21, npoints=npoints, is_weighting=True, **flags)
            np.random.seed(seed)
            ds.generate_chunks()
            # copied BEFORE __getitem__: with neither colour, normal nor multiview columns the reference's
            # point_set is a view of chunk_data (lib/dataset.py:68) and _augment writes through it
            out[name + "/chunks"] = np.stack([ds.chunk_data[k].copy() for k in sorted(sc)])
            items = [ds[i] for i in range(len(ds))]
            coords, feats, sem, w, _ = ds_mod.collate_random(items)
            out[name + "/labelweights"] = np.asarray(ds.labelweights)
            out[name + "/coords"] = coords.numpy()
            out[name + "/feats"] = feats.numpy()
            out[name + "/labels"] = sem.numpy()
            out[name + "/weights"] = w.numpy()
        for name, s, n, npoints, flags, seed in SCENE_CASES:
            sc = scenes_for([s], n, flags.get("use_multiview", False))
            for k, (scene, _) in sc.items():
                np.save(os.path.join(tmp, k + ".npy"), scene)
            ds_mod = load_reference_dataset(tmp, {k: f for k, (_, f) in sc.items()})
            ds = ds_mod.ScannetDatasetWholeScene(sorted(sc), npoints=npoints, is_weighting=True, **flags)
            np.random.seed(seed)
            pts, sem, w, _ = ds[0]
            out[name + "/labelweights"] = np.asarray(ds.labelweights)
            out[name + "/points"] = pts.astype(np.float32)   # collate_wholescene's FloatTensor cast
            out[name + "/labels"] = sem.astype(np.int64)
            out[name + "/weights"] = w.astype(np.float32)
    path = os.path.join(HERE, "ref_chunker.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB;", len(out), "arrays")


if __name__ == "__main__":
    main()
