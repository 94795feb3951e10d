"""oracle/chunker_np.py (NumPy restatement of lib/dataset.py's chunk generator) against the golden
outputs of the reference's own classes (tests/golden/make_chunker_golden.py, run on /root/reference)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_chunker_golden as G  # noqa: E402  (case tables and the scene generator only)
from oracle import chunker_np as O  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "ref_chunker.npz"))


def run_train_case(seeds, n, npoints, flags, seed, train):
    sc = G.scenes_for(seeds, n, flags.get("use_multiview", False))
    names = sorted(sc)
    lw = O.label_weights([sc[k][0] for k in names], 21)
    np.random.seed(seed)
    chunks = [O.generate_chunk(sc[k][0], npoints, np.random, feature=sc[k][1])[0] for k in names]
    items = [O.get_item(c, lw, train, rng=np.random, **flags)[:3] for c in chunks]
    return lw, np.stack(chunks), O.collate(items)


@pytest.mark.parametrize("case", G.TRAIN_CASES, ids=[c[0] for c in G.TRAIN_CASES])
def test_train_chunks_equal_reference(golden, case):
    name, seeds, n, npoints, flags, seed = case
    lw, chunks, (coords, feats, lab, w) = run_train_case(seeds, n, npoints, flags, seed, not name.startswith("val"))
    assert np.array_equal(lw, golden[name + "/labelweights"])
    assert np.array_equal(chunks, golden[name + "/chunks"])          # same rows drawn: same RNG stream
    assert np.array_equal(lab, golden[name + "/labels"])
    assert np.array_equal(w, golden[name + "/weights"])
    assert np.array_equal(feats, golden[name + "/feats"])
    # float64 augmentation restated op by op: equal after the float32 cast
    assert np.array_equal(coords, golden[name + "/coords"])


@pytest.mark.parametrize("case", G.SCENE_CASES, ids=[c[0] for c in G.SCENE_CASES])
def test_whole_scene_tiles_equal_reference(golden, case):
    name, s, n, npoints, flags, seed = case
    sc = G.scenes_for([s], n, flags.get("use_multiview", False))
    scene, feat = sc[sorted(sc)[0]]
    lw = O.label_weights([scene], 21)
    np.random.seed(seed)
    kw = {k: v for k, v in flags.items() if k != "use_multiview"}
    pts, lab, w, _ = O.whole_scene_tiles(scene, lw, npoints, feature=feat, rng=np.random, **kw)
    assert np.array_equal(lw, golden[name + "/labelweights"])
    assert np.array_equal(pts.astype(np.float32), golden[name + "/points"])
    assert np.array_equal(lab.astype(np.int64), golden[name + "/labels"])
    assert np.array_equal(w.astype(np.float32), golden[name + "/weights"])
