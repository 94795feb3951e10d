"""CUDA chunk generator (csrc/chunker.cu behind pointnet2/scannet_b200/chunker.py) against the golden
outputs of the reference's own dataset classes and against the NumPy oracle, through the C ABI.

Bar: crop selections, voxel / label statistics, drawn rows, labels, weights and every un-augmented
column bit-exact; augmented coordinates equal after the float32 cast except where NumPy's BLAS matmul
or sequential mean rounds the 53rd bit differently: |diff| <= 1 float32 ulp of the coordinate
(9.6e-7 at 8 m) on at most 0.1 % of the entries."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_chunker_golden as G  # noqa: E402
from oracle import chunker_np as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "ref_chunker.npz"))


def coords_close(got, want):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    ulp = np.spacing(np.abs(want)).astype(np.float64)
    assert (diff <= ulp).all(), "coordinate off by more than one float32 ulp: %g" % diff.max()
    assert (diff > 0).mean() <= 1e-3, "too many last-bit differences: %g" % (diff > 0).mean()


@pytest.mark.parametrize("case", G.TRAIN_CASES, ids=[c[0] for c in G.TRAIN_CASES])
def test_train_batches_equal_reference(golden, case):
    from pointnet2.scannet_b200 import chunker
    name, seeds, n, npoints, flags, seed = case
    sc = G.scenes_for(seeds, n, flags.get("use_multiview", False))
    names = sorted(sc)
    ds = chunker.ScannetDataset("val" if name.startswith("val") else "train", names,
                                {k: sc[k][0] for k in names}, num_classes=21, npoints=npoints,
                                multiview_data={k: sc[k][1] for k in names}, **flags)
    assert np.array_equal(ds.labelweights, golden[name + "/labelweights"])
    np.random.seed(seed)
    ds.generate_chunks()
    want_chunks = golden[name + "/chunks"]
    for b, k in enumerate(names):                       # the same scene rows were drawn
        rows = ds.chunk_data[k].cpu().numpy()
        assert np.array_equal(sc[k][0][rows], want_chunks[b][:, :11])
    coords, feats, lab, w, _ = ds.batch(range(len(ds)))
    assert lab.dtype == torch.int64 and coords.dtype == torch.float32
    assert np.array_equal(lab.cpu().numpy(), golden[name + "/labels"])
    assert np.array_equal(w.cpu().numpy(), golden[name + "/weights"])
    assert np.array_equal(feats.cpu().numpy(), golden[name + "/feats"])
    coords_close(coords.cpu().numpy(), golden[name + "/coords"])


def test_crop_statistics_equal_oracle():
    """Every try of generate_chunks: row count, annotated rows, occupied voxels, validity."""
    from pointnet2.scannet_b200 import chunker
    scenes = {"s%d" % s: G.synthetic.make_prep_scene(s, npoints=25000) for s in (31, 32, 33)}
    ds = chunker.ScannetDataset("train", sorted(scenes), scenes, npoints=512, use_color=True)
    np.random.seed(7)
    ds.generate_chunks()
    np.random.seed(7)
    for k in sorted(scenes):
        _, tr = O.generate_chunk(scenes[k], 512, np.random)
        assert [tuple(t) for t in tr["tries"]] == [tuple(t) for t in ds.tries[k]]
        assert np.array_equal(tr["rows"][tr["choices"]], ds.chunk_data[k].cpu().numpy())


@pytest.mark.parametrize("case", G.SCENE_CASES, ids=[c[0] for c in G.SCENE_CASES])
def test_whole_scene_tiles_equal_reference(golden, case):
    from pointnet2.scannet_b200 import chunker
    name, s, n, npoints, flags, seed = case
    sc = G.scenes_for([s], n, flags.get("use_multiview", False))
    k = sorted(sc)[0]
    ds = chunker.ScannetDatasetWholeScene([k], {k: sc[k][0]}, npoints=npoints, multiview_data={k: sc[k][1]},
                                          **flags)
    np.random.seed(seed)
    pts, lab, w, _ = ds[0]
    assert np.array_equal(pts.cpu().numpy(), golden[name + "/points"])   # no arithmetic but rgb / 255
    assert np.array_equal(lab.cpu().numpy(), golden[name + "/labels"])
    assert np.array_equal(w.cpu().numpy(), golden[name + "/weights"])


def test_full_size_scene_and_box_properties():
    """BASELINE-sized input: a 150k-point scene, 8192-point chunks and tiles, against the oracle;
    size-independent properties of the selection: ascending, complete, inside the box."""
    from pointnet2.scannet_b200 import chunker, lib
    scene = G.synthetic.make_prep_scene(5, npoints=150000, size=(6.0, 5.0, 2.7))
    ds = chunker.ScannetDataset("train", ["a"], {"a": scene}, npoints=8192, use_color=True, use_normal=True)
    np.random.seed(3)
    ds.generate_chunks()
    x, lab, w, _ = ds[0]
    np.random.seed(3)
    lw = O.label_weights([scene], 21)
    chunk, tr = O.generate_chunk(scene, 8192, np.random)
    ps, l, sw, _ = O.get_item(chunk, lw, True, use_color=True, use_normal=True, rng=np.random)
    assert np.array_equal(lab.cpu().numpy(), l.astype(np.int64))
    assert np.array_equal(w.cpu().numpy(), sw.astype(np.float32))
    assert np.array_equal(x[:, 3:].cpu().numpy(), ps[:, 3:].astype(np.float32))
    coords_close(x[:, :3].cpu().numpy(), ps[:, :3].astype(np.float32))
    # selection properties on the last try's box
    st = ds.stores["a"]
    centre_row, count = ds.tries["a"][-1][:2]
    sel = st.sel[:count].cpu().numpy()
    assert (np.diff(sel) > 0).all()
    c = scene[centre_row, :3]
    lo = np.array([c[0] - 0.75, c[1] - 0.75, st.coordmin[2]]) - 0.2
    hi = np.array([c[0] + 0.75, c[1] + 0.75, st.coordmax[2]]) + 0.2
    inside = np.all((scene[:, :3] >= lo) & (scene[:, :3] <= hi), axis=1)
    assert np.array_equal(np.nonzero(inside)[0], sel)
    # whole-scene tiling at full size
    ws = chunker.ScannetDatasetWholeScene(["a"], {"a": scene}, npoints=8192, use_color=True, use_normal=True)
    np.random.seed(4)
    pts, lab, w, _ = ws[0]
    np.random.seed(4)
    op, ol, ow, _ = O.whole_scene_tiles(scene, lw, 8192, use_color=True, use_normal=True, rng=np.random)
    assert np.array_equal(pts.cpu().numpy(), op.astype(np.float32))
    assert np.array_equal(lab.cpu().numpy(), ol.astype(np.int64))
    assert np.array_equal(w.cpu().numpy(), ow.astype(np.float32))
    assert lib.launch_count() > 0


def test_rejects_unrepresentable_scene_and_bad_arguments():
    from pointnet2.scannet_b200 import chunker, lib
    scene = G.synthetic.make_prep_scene(6, npoints=2000)
    bad = scene.copy()
    bad[0, 0] += 1e-12
    with pytest.raises(ValueError):
        chunker.SceneStore(bad)
    L = lib.load()
    assert L.pn2_crop_try(0, 11, None, 10, 0, 0.0, 1.0, None, None, None) == -1
    st = chunker.SceneStore(scene)
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert L.pn2_crop_try(st.n, st.ld, st.rows.data_ptr(), 10, st.n, 0.0, 1.0, st.sel.data_ptr(),
                          stats.data_ptr(), None) == -1          # centre row out of range
