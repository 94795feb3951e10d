"""NumPy oracle of the reference's chunk generator -- TEST INFRASTRUCTURE ONLY.

Restates, step by step and in the reference's own dtypes (float64 scenes, see
/root/reference/preprocessing/collect_scannet_scenes.py:64-79: float32 points concatenated
with float64 label columns), what feeds the SA/FP hot path:

* ``crop_candidates``   lib/dataset.py:220-237   one try of ScannetDataset.generate_chunks
* ``generate_chunk``    lib/dataset.py:216-249   the five tries + the draw with replacement
* ``augment``           lib/dataset.py:99-200    _augment / _translate / _rotate / _scale
* ``get_item``          lib/dataset.py:61-94     ScannetDataset.__getitem__ (column assembly, weights)
* ``whole_scene_tiles`` lib/dataset.py:293-353   ScannetDatasetWholeScene.__getitem__
* ``collate``           lib/dataset.py:358-396   collate_random (float32 / int64 casts, coords | feats)

Randomness is drawn from the ``rng`` argument with exactly the calls, in exactly the order, of the
reference (``np.random.choice`` on the same populations), so seeding ``np.random`` identically and
passing ``rng=np.random`` reproduces the reference's chunks; tests/golden/make_chunker_golden.py
pins this file against the reference's own classes (imported from /root/reference with its missing
third-party imports stubbed).  Only tests/ may import this module; the product
(pointnet2/scannet_b200/chunker.py) never does.
"""
import numpy as np

XYZ_HALF = np.array([0.75, 0.75, 1.5])  # lib/dataset.py:222-223
VOXELS = np.array([31.0, 31.0, 62.0])   # lib/dataset.py:237


def label_weights(scenes, num_classes, is_weighting=True):
    """lib/dataset.py:49-58 (float32 arithmetic after the astype)."""
    if not is_weighting:
        return np.ones(num_classes)
    hist = np.zeros(num_classes)
    for scene in scenes:
        h, _ = np.histogram(scene[:, 10], range(num_classes + 1))
        hist += h
    hist = hist.astype(np.float32)
    hist = hist / np.sum(hist)
    return 1 / np.log(1.2 + hist)


def crop_candidates(scene, centre_row, coordmin, coordmax):
    """One try (lib/dataset.py:220-238).  Returns (selected row indices, curmin, curmax, isvalid,
    n_labelled, n_voxels); isvalid is None when nothing was selected (the reference `continue`s)."""
    curcenter = scene[centre_row, :3]
    curmin = curcenter - XYZ_HALF
    curmax = curcenter + XYZ_HALF
    curmin[2] = coordmin[2]
    curmax[2] = coordmax[2]
    xyz = scene[:, :3]
    choice = np.sum((xyz >= (curmin - 0.2)) * (xyz <= (curmax + 0.2)), axis=1) == 3
    rows = np.nonzero(choice)[0]
    if len(rows) == 0:
        return rows, curmin, curmax, None, 0, 0
    cur = scene[rows]
    sem = cur[:, 10].astype(np.int32)
    inner = np.sum((cur[:, :3] >= (curmin - 0.01)) * (cur[:, :3] <= (curmax + 0.01)), axis=1) == 3
    vidx = np.ceil((cur[inner, :3] - curmin) / (curmax - curmin) * VOXELS)
    keys = np.unique(vidx[:, 0] * 31.0 * 62.0 + vidx[:, 1] * 62.0 + vidx[:, 2])
    n_lab = int(np.sum(sem > 0))
    isvalid = bool(n_lab / len(sem) >= 0.7 and len(keys) / 31.0 / 31.0 / 62.0 >= 0.02)
    return rows, curmin, curmax, isvalid, n_lab, len(keys)


def generate_chunk(scene, npoints, rng=np.random, feature=None):
    """lib/dataset.py:209-249 for one scene: (chunk (npoints, 11 [+ F]) float64, trace)."""
    coordmax = np.max(scene, axis=0)[:3]
    coordmin = np.min(scene, axis=0)[:3]
    rows = None
    trace = []
    for _ in range(5):
        centre_row = rng.choice(len(scene), 1)[0]
        r, curmin, curmax, isvalid, n_lab, n_vox = crop_candidates(scene, centre_row, coordmin, coordmax)
        trace.append((int(centre_row), len(r), n_lab, n_vox, isvalid))
        if isvalid is None:
            continue  # keeps the previous try's selection, like the reference's stale variables
        rows = r
        if isvalid:
            break
    cur = scene[rows]
    if feature is not None:
        cur = np.concatenate([cur, feature[rows]], axis=1)
    choices = rng.choice(cur.shape[0], npoints, replace=True)
    return cur[choices], {"rows": rows, "choices": choices, "tries": trace}


def draw_augmentation(rng=np.random):
    """The random draws of _augment / _translate / _rotate / _scale in the reference's order
    (lib/dataset.py:104-137, 148-150, 162-186, 200): (ops, t[3], R[3, 3], f)."""
    p = rng.choice(np.arange(0.01, 1.01, 0.01), size=1)[0]
    if p < 1 / 8:
        ops = "t"
    elif p < 2 / 8:
        ops = "r"
    elif p < 3 / 8:
        ops = "s"
    elif p < 4 / 8:
        ops = "tr"
    elif p < 5 / 8:
        ops = "ts"
    elif p < 6 / 8:
        ops = "rs"
    elif p < 7 / 8:
        ops = "trs"
    else:
        ops = ""
    t, R, f = np.zeros(3), np.eye(3), 1.0
    for op in ops:
        if op == "t":
            t = np.array([rng.choice(np.arange(-0.5, 0.501, 0.001), size=1)[0] for _ in range(3)])
        elif op == "r":
            th = [rng.choice(np.arange(-5, 5.001, 0.001), size=1)[0] * 3.14 / 180 for _ in range(3)]
            Rx = np.array([[1, 0, 0], [0, np.cos(th[0]), -np.sin(th[0])], [0, np.sin(th[0]), np.cos(th[0])]])
            Ry = np.array([[np.cos(th[1]), 0, np.sin(th[1])], [0, 1, 0], [-np.sin(th[1]), 0, np.cos(th[1])]])
            Rz = np.array([[np.cos(th[2]), -np.sin(th[2]), 0], [np.sin(th[2]), np.cos(th[2]), 0], [0, 0, 1]])
            R = np.matmul(np.matmul(Rz, Ry), Rx)
        else:
            f = rng.choice(np.arange(0.95, 1.051, 0.001), size=1)[0]
    return ops, t, R, f


def apply_augmentation(xyz, ops, t, R, f):
    """lib/dataset.py:101-102, 139-141: about the chunk's mean, in float64."""
    centre = np.mean(xyz, axis=0)
    c = xyz - centre
    for op in ops:
        if op == "t":
            c = c + t
        elif op == "r":
            c = np.matmul(R, c.T).T
        else:
            c = c * np.array([f, f, f])
    return c + centre


def get_item(chunk, labelweights, train, use_multiview=False, use_color=False, use_normal=False,
             rng=np.random):
    """ScannetDataset.__getitem__ (lib/dataset.py:61-94) on one stored chunk:
    (point_set (npoints, 3 + C) float64, label int32, sample_weight)."""
    cols = [chunk[:, :3]]
    if use_multiview:
        cols.append(chunk[:, 11:])
    if use_color:
        cols.append(chunk[:, 3:6] / 255.)
    if use_normal:
        cols.append(chunk[:, 6:9])
    point_set = np.concatenate(cols, axis=1)
    label = chunk[:, 10].astype(np.int32)
    aug = None
    if train:
        aug = draw_augmentation(rng)
        point_set[:, :3] = apply_augmentation(point_set[:, :3], *aug)
    curmin = np.min(point_set, axis=0)[:3]
    curmax = np.max(point_set, axis=0)[:3]
    mask = np.sum((point_set[:, :3] >= (curmin - 0.01)) * (point_set[:, :3] <= (curmax + 0.01)), axis=1) == 3
    sample_weight = labelweights[label] * mask
    return point_set, label, sample_weight, aug


def whole_scene_tiles(scene, labelweights, npoints, use_color=False, use_normal=False, feature=None,
                      rng=np.random):
    """ScannetDatasetWholeScene.__getitem__ (lib/dataset.py:293-353).  Note the channel order
    (multiview LAST, :306-308) and `mask = mask[choice]` (:334): the scene-length box mask indexed
    with selection-relative indices -- reproduced as it is."""
    cols = [scene[:, :3]]
    if use_color:
        cols.append(scene[:, 3:6] / 255.)
    if use_normal:
        cols.append(scene[:, 6:9])
    if feature is not None:
        cols.append(feature)
    pts = np.concatenate(cols, axis=1)
    sem = scene[:, 10].astype(np.int32)
    coordmax = pts[:, :3].max(axis=0)
    coordmin = pts[:, :3].min(axis=0)
    nx = np.ceil((coordmax[0] - coordmin[0]) / 1.5).astype(np.int32)
    ny = np.ceil((coordmax[1] - coordmin[1]) / 1.5).astype(np.int32)
    out_p, out_l, out_w, trace = [], [], [], []
    for i in range(nx):
        for j in range(ny):
            curmin = coordmin + [i * 1.5, j * 1.5, 0]
            curmax = coordmin + [(i + 1) * 1.5, (j + 1) * 1.5, coordmax[2] - coordmin[2]]
            mask = np.sum((pts[:, :3] >= (curmin - 0.01)) * (pts[:, :3] <= (curmax + 0.01)), axis=1) == 3
            rows = np.nonzero(mask)[0]
            if len(rows) == 0:
                continue
            choice = rng.choice(len(rows), npoints, replace=True)
            out_p.append(pts[rows[choice]])
            out_l.append(sem[rows[choice]])
            out_w.append(labelweights[sem[rows[choice]]] * mask[choice])
            trace.append((i, j, len(rows)))
    return np.stack(out_p), np.stack(out_l), np.stack(out_w), trace


def collate(items):
    """collate_random (lib/dataset.py:358-396): float32 points, int64 labels, float32 weights."""
    pts = np.stack([it[0] for it in items]).astype(np.float32)
    lab = np.stack([it[1] for it in items]).astype(np.int64)
    w = np.stack([it[2] for it in items]).astype(np.float32)
    return pts[:, :, :3], pts[:, :, 3:], lab, w
