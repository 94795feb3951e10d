"""Synthetic ScanNet-shaped scene chunks (no dataset is available offline).

Mirrors what the reference feeds the op path:

* ``ScannetDataset.generate_chunks`` (/root/reference/lib/dataset.py:204-250):
  a 1.5 m x 1.5 m full-height column (+0.2 m margin) cut from a scene, then
  ``npoints`` indices drawn WITH replacement (:248) -- so a chunk holds many
  exact duplicate points, and exact distance ties are the common case.
* ``ScannetDataset.__getitem__`` (:60-94): channel order
  xyz, [multiview 128], [rgb in 0..1], [normal]; coordinates are absolute scene
  coordinates (not centred).
* ``ScannetDatasetWholeScene.__getitem__`` (:293-353): tile the scene's XY
  bounding box into 1.5 m columns, resample each to ``npoints``.

Everything is numpy on the host, seeded, and returns float32 arrays.
"""
import numpy as np

NUM_CLASSES = 20  # /root/reference/lib/config.py NYU-20 label set


def _plane(rng, origin, u, v, step=0.02):
    nu = max(int(np.linalg.norm(u) / step), 1)
    nv = max(int(np.linalg.norm(v) / step), 1)
    a, b = np.meshgrid(np.arange(nu) / nu, np.arange(nv) / nv, indexing="ij")
    pts = origin + a.reshape(-1, 1) * u + b.reshape(-1, 1) * v
    nrm = np.cross(u, v)
    nrm = nrm / (np.linalg.norm(nrm) + 1e-12)
    return pts, np.broadcast_to(nrm, pts.shape)


def surface_pool(rng, extent=(1.9, 1.9, 2.7), n_boxes=None):
    """Surface samples (floor, two walls, a few boxes) of one column: (P,3), (P,3)."""
    ex, ey, ez = extent
    parts = [
        _plane(rng, np.zeros(3), np.array([ex, 0, 0.0]), np.array([0, ey, 0.0])),
        _plane(rng, np.zeros(3), np.array([ex, 0, 0.0]), np.array([0, 0, ez])),
        _plane(rng, np.zeros(3), np.array([0, ey, 0.0]), np.array([0, 0, ez])),
    ]
    n_boxes = int(rng.integers(3, 6)) if n_boxes is None else n_boxes
    for _ in range(n_boxes):
        size = rng.uniform([0.3, 0.3, 0.3], [0.9, 0.9, 1.2])
        org = rng.uniform([0.1, 0.1, 0.0], [ex - size[0], ey - size[1], 0.05])
        sx, sy, sz = size
        parts.append(_plane(rng, org + [0, 0, sz], np.array([sx, 0, 0.0]), np.array([0, sy, 0.0])))
        parts.append(_plane(rng, org, np.array([sx, 0, 0.0]), np.array([0, 0, sz])))
        parts.append(_plane(rng, org, np.array([0, sy, 0.0]), np.array([0, 0, sz])))
        parts.append(_plane(rng, org + [sx, 0, 0], np.array([0, sy, 0.0]), np.array([0, 0, sz])))
        parts.append(_plane(rng, org + [0, sy, 0], np.array([sx, 0, 0.0]), np.array([0, 0, sz])))
    pts = np.concatenate([p for p, _ in parts])
    nrm = np.concatenate([n for _, n in parts])
    pts = pts + rng.normal(0.0, 0.003, pts.shape)
    return pts, nrm


def make_chunk(seed, npoints=8192, use_color=True, use_normal=True, use_multiview=False,
               pool_size=20000):
    """One training chunk: (npoints, 3 + C) float32, labels (npoints,) int64."""
    rng = np.random.default_rng(seed)
    pts, nrm = surface_pool(rng)
    if len(pts) > pool_size:
        keep = rng.choice(len(pts), pool_size, replace=False)
        pts, nrm = pts[keep], nrm[keep]
    pts = pts + rng.uniform(0.5, 7.0, 3) * [1, 1, 0]  # absolute scene coordinates
    cols = [pts]
    if use_multiview:
        feat = np.maximum(rng.normal(0, 1, (len(pts), 128)), 0)
        feat[rng.random(len(pts)) < 0.3] = 0.0  # points no camera saw
        cols.append(feat)
    if use_color:
        cols.append(rng.random((len(pts), 3)))
    if use_normal:
        cols.append(nrm)
    pool = np.concatenate(cols, axis=1)
    pick = rng.choice(len(pool), npoints, replace=True)  # dataset.py:248
    labels = rng.integers(0, NUM_CLASSES, len(pool))[pick]
    return pool[pick].astype(np.float32), labels.astype(np.int64)


def make_batch(config_id, batch, npoints=8192, first_chunk=0, **kw):
    """(batch, npoints, 3+C) float32, (batch, npoints) int64, weights (batch, npoints) float32.

    Chunk ``i`` uses ``numpy.random.default_rng(1000 * config_id + first_chunk + i)``.
    """
    xs, ys = [], []
    for i in range(batch):
        x, y = make_chunk(1000 * config_id + first_chunk + i, npoints, **kw)
        xs.append(x), ys.append(y)
    x = np.stack(xs)
    y = np.stack(ys)
    return x, y, np.ones(y.shape, np.float32)


def input_channels(use_color=True, use_normal=True, use_multiview=False):
    """scripts/train.py:39"""
    return 3 * int(use_color) + 3 * int(use_normal) + 128 * int(use_multiview)


def make_scene(seed, npoints=150000, size=(6.0, 5.0, 2.7)):
    """A whole synthetic room: (npoints, 9) float32 [xyz rgb normal]."""
    rng = np.random.default_rng(seed)
    sx, sy, sz = size
    tiles = []
    for ix in range(int(np.ceil(sx / 1.9))):
        for iy in range(int(np.ceil(sy / 1.9))):
            p, n = surface_pool(rng)
            p = p + [ix * 1.9, iy * 1.9, 0.0]
            tiles.append(np.concatenate([p, rng.random((len(p), 3)), n], axis=1))
    pool = np.concatenate(tiles)
    pool = pool[(pool[:, 0] <= sx) & (pool[:, 1] <= sy)]
    pick = rng.choice(len(pool), npoints, replace=len(pool) < npoints)
    scene = pool[pick]
    scene[:, :3] += [0.7, 1.1, 0.0]
    return scene.astype(np.float32)


def tile_scene(scene, npoints=8192, seed=0):
    """Whole-scene column tiling of lib/dataset.py:316-345 -> (K, npoints, C) float32."""
    rng = np.random.default_rng(seed)
    cmax, cmin = scene[:, :3].max(0), scene[:, :3].min(0)
    nx = int(np.ceil((cmax[0] - cmin[0]) / 1.5))
    ny = int(np.ceil((cmax[1] - cmin[1]) / 1.5))
    out = []
    for i in range(nx):
        for j in range(ny):
            lo = cmin + [i * 1.5, j * 1.5, 0]
            hi = cmin + [(i + 1) * 1.5, (j + 1) * 1.5, cmax[2] - cmin[2]]
            m = np.all((scene[:, :3] >= lo - 0.01) & (scene[:, :3] <= hi + 0.01), axis=1)
            if not m.any():
                continue
            sub = scene[m]
            out.append(sub[rng.choice(len(sub), npoints, replace=True)])
    return np.stack(out).astype(np.float32)


def adversarial_clouds(seed=7):
    """Small clouds that hit the reference's corner cases."""
    rng = np.random.default_rng(seed)
    out = {}
    base = rng.uniform(0.5, 2.0, (100, 3)).astype(np.float32)
    dup = base[rng.integers(0, 20, 100)]  # heavy duplication -> exact ties
    out["duplicates_100"] = dup
    near = base.copy()
    near[0] = [0.01, 0.01, 0.01]  # |p|^2 <= 1e-3: never a candidate, yet idx[0] = 0
    near[17] = [0.0, 0.02, 0.0]
    near[63] = [0.03, 0.0, 0.0099]
    out["near_origin_100"] = near
    out["all_same_37"] = np.broadcast_to(base[:1], (37, 3)).copy()
    out["all_near_origin_40"] = (rng.uniform(-0.01, 0.01, (40, 3))).astype(np.float32)
    out["grid_ties_64"] = np.stack(np.meshgrid(*[np.arange(4.0)] * 3, indexing="ij"), -1) \
        .reshape(-1, 3).astype(np.float32) + 1.0
    out["n_300"] = rng.uniform(0.5, 2.0, (300, 3)).astype(np.float32)[rng.integers(0, 150, 300)]
    return out


def make_prep_scene(seed, npoints=40000, size=(4.0, 3.5, 2.7), num_classes=21, multiview=0):
    """A synthetic scene in the reference's PREPROCESSED file format
    (/root/reference/preprocessing/collect_scannet_scenes.py:64-79): float64 (npoints, 11) rows
    [x y z | r g b in 0..255 | nx ny nz | instance | semantic label], the point columns holding
    float32-representable values (they come from a float32 .ply).  About 15 % of the points are
    unannotated (label 0).  `multiview` > 0 also returns a float32 (npoints, multiview) feature
    matrix (the h5py rows of lib/dataset.py:46-47)."""
    rng = np.random.default_rng(seed)
    s9 = make_scene(seed, npoints=npoints, size=size).astype(np.float64)
    s9[:, 3:6] = np.floor(s9[:, 3:6] * 256.0).clip(0, 255)
    label = rng.integers(1, num_classes, npoints).astype(np.float64)
    label[rng.random(npoints) < 0.15] = 0.0
    inst = rng.integers(0, 40, npoints).astype(np.float64)
    scene = np.concatenate([s9, inst[:, None], label[:, None]], axis=1)
    if multiview:
        feat = np.maximum(rng.normal(0, 1, (npoints, multiview)), 0).astype(np.float32)
        feat[rng.random(npoints) < 0.3] = 0.0
        return scene, feat
    return scene
