"""Chunk generation with the scenes resident in HBM -- mirror of the reference's datasets
(/root/reference/lib/dataset.py), served by csrc/chunker.cu.

``ScannetDataset`` / ``ScannetDatasetWholeScene`` / ``collate_random`` keep the reference's
names, constructor flags, method names and outputs; what differs is where the data lives:

* the reference ``np.load``s every preprocessed scene (``CONF.SCANNETV2_FILE``, float64 ``(N, 11)``
  rows ``x y z | r g b | nx ny nz | instance | label``) and crops / resamples / augments it with
  NumPy on one host core (``num_workers=0``, scripts/train.py:115); here the caller hands the same
  arrays over once (``scene_data`` dict) and they stay on the GPU as float32 rows -- the file
  format's values are float32-representable, which ``SceneStore`` verifies (no silent rounding);
* every crop test, voxel key, rgb / 255 and augmentation product is evaluated in double on the
  device, in NumPy's order, so selections, labels, weights and un-augmented columns are
  bit-identical to the reference and augmented coordinates agree to the last float32 bit except
  where NumPy's own BLAS / summation order rounds the 53rd bit differently;
* randomness is drawn from ``np.random`` with exactly the reference's calls in the reference's
  order (so ``np.random.seed(s)`` gives the reference's chunks).  The draws depend on the number
  of rows in the crop, hence one small device->host read per try -- the price of RNG parity.

Outputs are CUDA tensors in ``collate_random``'s dtypes: ``__getitem__`` returns
``(point_set (npoints, 3 + C) float32, label int64, sample_weight float32, fetch_time)``;
``batch(indices)`` assembles straight into ``(B, npoints, 3 + C)`` and returns
``(coords, feats, semantic_segs, sample_weights, fetch_time)``.

There is no CPU path: the library is CUDA-only (lib.load() raises if it is missing).
"""
import ctypes
import time

import numpy as np
import torch

from . import lib as _lib

LABEL_COL = 10  # lib/dataset.py:39, 71
# one record per chunk of a batched launch: csrc/chunker.cu's ChunkDesc (size checked at first use)
DESC = np.dtype([("scene", "u8"), ("feature", "u8"), ("sel", "u8"), ("choice", "u8"), ("wmask", "u8"),
                 ("ld", "i4"), ("nfeat", "i4"), ("ops", "i4"), ("_pad", "i4"), ("t", "f8", 3), ("R", "f8", 9),
                 ("f", "f8"), ("centre", "f8", 3)])


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class SceneStore(object):
    """One preprocessed scene in HBM: float32 rows [n, ld] (+ optional multiview rows [n, F])."""

    def __init__(self, scene, feature=None, device="cuda"):
        scene = np.asarray(scene)
        if scene.ndim != 2 or scene.shape[1] < 11:
            raise ValueError("scene must be (N, >= 11): x y z r g b nx ny nz instance label")
        rows32 = np.ascontiguousarray(scene, dtype=np.float32)
        if not np.array_equal(rows32.astype(scene.dtype), scene):
            raise ValueError("scene values are not float32-representable; the device copy would round them")
        self.n, self.ld = rows32.shape
        self.device = torch.device(device)
        self.rows = torch.from_numpy(rows32).to(self.device)
        # np.max / np.min over the float64 scene (lib/dataset.py:216-217, 312-313)
        self.coordmax = scene[:, :3].max(axis=0).astype(np.float64)
        self.coordmin = scene[:, :3].min(axis=0).astype(np.float64)
        self.feature = None
        self.nfeat = 0
        if feature is not None:
            f32 = np.ascontiguousarray(feature, dtype=np.float32)
            if f32.shape[0] != self.n or not np.array_equal(f32.astype(np.asarray(feature).dtype), feature):
                raise ValueError("multiview features must be (N, F) and float32-representable")
            self.feature = torch.from_numpy(f32).to(self.device)
            self.nfeat = f32.shape[1]
        self.sel = torch.empty(self.n, dtype=torch.int32, device=self.device)  # rows of the current crop
        self.count = 0


def _label_weights(scenes, num_classes, is_weighting):
    """lib/dataset.py:49-58 / :277-286 (one-off host arithmetic, in the reference's dtypes)."""
    if not is_weighting:
        return np.ones(num_classes)
    hist = np.zeros(num_classes)
    for s in scenes:
        h, _ = np.histogram(s[:, LABEL_COL], range(num_classes + 1))
        hist += h
    hist = hist.astype(np.float32)
    hist = hist / np.sum(hist)
    return 1 / np.log(1.2 + hist)


# the populations of the reference's np.random.choice calls (lib/dataset.py:104, 148-150, 162-186, 200);
# choice() on a cached array consumes the generator exactly like choice() on a fresh np.arange
_POP_P = np.arange(0.01, 1.01, 0.01)
_POP_T = np.arange(-0.5, 0.501, 0.001)
_POP_R = np.arange(-5, 5.001, 0.001)
_POP_S = np.arange(0.95, 1.051, 0.001)


def _draw_augmentation():
    """The draws of _augment / _translate / _rotate / _scale (lib/dataset.py:104-137, 148-150, 162-186,
    200) from np.random, in the reference's order -> (ops bit mask, 13 doubles t | R | f)."""
    p = np.random.choice(_POP_P, size=1)[0]
    eighth = int(p >= 1 / 8) + int(p >= 2 / 8) + int(p >= 3 / 8) + int(p >= 4 / 8) + int(p >= 5 / 8) \
        + int(p >= 6 / 8) + int(p >= 7 / 8)
    ops = ("t", "r", "s", "tr", "ts", "rs", "trs", "")[eighth]
    params = np.zeros(13)
    params[3:12] = np.eye(3).ravel()
    params[12] = 1.0
    mask = 0
    for op in ops:
        if op == "t":
            mask |= 1
            for d in range(3):
                params[d] = np.random.choice(_POP_T, size=1)[0]
        elif op == "r":
            mask |= 2
            a, b, c = (np.random.choice(_POP_R, size=1)[0] * 3.14 / 180 for _ in range(3))
            Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
            Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
            Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
            params[3:12] = np.matmul(np.matmul(Rz, Ry), Rx).ravel()
        else:
            mask |= 4
            params[12] = np.random.choice(_POP_S, size=1)[0]
    return mask, params


class _Base(object):
    def _setup(self, scene_list, scene_data, multiview_data, num_classes, is_weighting, device):
        if not torch.cuda.is_available():
            raise RuntimeError("the chunk generator needs a CUDA device (no CPU path)")
        self._L = _lib.load()
        self.device = torch.device(device)
        self.scene_list = list(scene_list)
        feats = multiview_data or {}
        self.stores = {k: SceneStore(scene_data[k], feats.get(k) if self.use_multiview else None, device)
                       for k in self.scene_list}
        self.labelweights = _label_weights([scene_data[k] for k in self.scene_list], num_classes, is_weighting)
        self._lw = torch.from_numpy(np.ascontiguousarray(self.labelweights, dtype=np.float32)).to(self.device)
        self._centre = torch.zeros(3, dtype=torch.float64, device=self.device)
        self.num_classes = num_classes

    def channels(self):
        return 3 * int(self.use_color) + 3 * int(self.use_normal) + \
            (next(iter(self.stores.values())).nfeat if self.use_multiview else 0)

    def _assemble_batch(self, entries, npoints, out, labels, weights, feat_first):
        """entries: [(store, sel or None, choice, wmask or None, aug or None)] -> out [len, npoints, ldo] in
        two launches (one descriptor table, one host->device copy)."""
        L, dev = self._L, self.device
        if L.pn2_chunk_desc_bytes() != DESC.itemsize:
            raise RuntimeError("ChunkDesc layout mismatch: rebuild libpn2_b200.so")
        desc = np.zeros(len(entries), DESC)
        for d, (st, sel, choice, wmask, aug) in zip(desc, entries):
            d["scene"], d["ld"], d["nfeat"] = st.rows.data_ptr(), st.ld, st.nfeat
            d["feature"] = st.feature.data_ptr() if st.nfeat else 0
            d["sel"] = sel.data_ptr() if sel is not None else 0
            d["choice"] = choice.data_ptr()
            d["wmask"] = wmask.data_ptr() if wmask is not None else 0
            if aug is not None and aug[0]:
                d["ops"], d["t"], d["R"], d["f"] = aug[0], aug[1][:3], aug[1][3:12], aug[1][12]
        table = torch.from_numpy(desc.view(np.uint8).reshape(len(entries), DESC.itemsize)).to(dev)
        with torch.cuda.device(dev):
            _lib.check(L.pn2_chunk_batch(len(entries), npoints, table.data_ptr(), int(feat_first),
                                         int(self.use_color), int(self.use_normal), LABEL_COL, self._lw.data_ptr(),
                                         self.num_classes, out.shape[-1], out.data_ptr(), labels.data_ptr(),
                                         weights.data_ptr(), _stream(dev)), "chunk_batch")

    def _assemble(self, st, sel, choice, out, labels, weights, feat_first, wmask=None, aug=None):
        L, dev = self._L, self.device
        npoints = choice.numel()
        ops, params = aug if aug is not None else (0, None)
        with torch.cuda.device(dev):
            if ops:
                _lib.check(L.pn2_chunk_centre(npoints, st.ld, st.rows.data_ptr(),
                                              sel.data_ptr() if sel is not None else None, choice.data_ptr(),
                                              self._centre.data_ptr(), _stream(dev)), "chunk_centre")
            host = params.ctypes.data_as(ctypes.c_void_p) if ops else None
            _lib.check(L.pn2_chunk_assemble(
                npoints, st.ld, st.rows.data_ptr(), sel.data_ptr() if sel is not None else None,
                choice.data_ptr(), st.feature.data_ptr() if st.nfeat else None, st.nfeat, int(feat_first),
                int(self.use_color), int(self.use_normal), LABEL_COL, self._lw.data_ptr(), self.num_classes,
                wmask.data_ptr() if wmask is not None else None, self._centre.data_ptr(), ops, host,
                out.shape[-1], out.data_ptr(), labels.data_ptr(), weights.data_ptr(), _stream(dev)),
                "chunk_assemble")


class ScannetDataset(_Base):
    """lib/dataset.py:14-250.  ``scene_data``: {scene_id: float64 (N, 11) array} (what the reference
    np.loads in _prepare_weights); ``multiview_data``: {scene_id: (N, F) features} when use_multiview."""

    def __init__(self, phase, scene_list, scene_data, num_classes=21, npoints=8192, is_weighting=True,
                 use_multiview=False, use_color=False, use_normal=False, multiview_data=None, device="cuda"):
        assert phase in ["train", "val", "test"]
        self.phase = phase
        self.npoints = npoints
        self.is_weighting = is_weighting
        self.use_multiview, self.use_color, self.use_normal = use_multiview, use_color, use_normal
        self._setup(scene_list, scene_data, multiview_data, num_classes, is_weighting, device)
        self.chunk_data = {}   # scene_id -> (selection rows, choices) on the device; see generate_chunks()
        self.tries = {}

    def __len__(self):
        return len(self.scene_list)

    def generate_chunks(self):
        """lib/dataset.py:204-250: per scene up to five crops around a random point, the first valid
        one (>= 70 % annotated rows, >= 2 % of the voxel grid occupied) is kept -- else the last --
        and ``npoints`` of its rows are drawn with replacement.  Stores row indices, not copies."""
        L, dev = self._L, self.device
        stats = torch.zeros(4, dtype=torch.int32, device=dev)
        for scene_id in self.scene_list:
            st = self.stores[scene_id]
            trace = []
            for _ in range(5):
                centre_row = int(np.random.choice(st.n, 1)[0])
                with torch.cuda.device(dev):
                    _lib.check(L.pn2_crop_try(st.n, st.ld, st.rows.data_ptr(), LABEL_COL, centre_row,
                                              float(st.coordmin[2]), float(st.coordmax[2]), st.sel.data_ptr(),
                                              stats.data_ptr(), _stream(dev)), "crop_try")
                count, n_lab, n_vox, n_odd = stats.tolist()   # the one device->host read of the try
                if n_odd:
                    raise RuntimeError("scene %s has zero extent along an axis of the crop box" % scene_id)
                # the centre point itself is inside its box, so count >= 1 (the reference's `continue`
                # on an empty crop, :232-233, is unreachable)
                isvalid = n_lab / count >= 0.7 and n_vox / 31.0 / 31.0 / 62.0 >= 0.02
                trace.append((centre_row, count, n_lab, n_vox, bool(isvalid)))
                if isvalid:
                    break
            choices = np.random.choice(count, self.npoints, replace=True)
            choice = torch.from_numpy(choices.astype(np.int32)).to(dev)
            rows = torch.gather(st.sel[:count], 0, choice.long())     # scene rows of the chunk
            self.chunk_data[scene_id] = rows.contiguous()
            self.tries[scene_id] = trace

    def _item_into(self, index, out, labels, weights):
        scene_id = self.scene_list[index]
        st = self.stores[scene_id]
        aug = _draw_augmentation() if self.phase == "train" else None
        # __getitem__'s mask (:86-88) compares the chunk with its own bounding box: always true
        self._assemble(st, None, self.chunk_data[scene_id], out, labels, weights, feat_first=True, aug=aug)

    def __getitem__(self, index):
        start = time.time()
        dev, c = self.device, 3 + self.channels()
        out = torch.empty(self.npoints, c, dtype=torch.float32, device=dev)
        labels = torch.empty(self.npoints, dtype=torch.int64, device=dev)
        weights = torch.empty(self.npoints, dtype=torch.float32, device=dev)
        self._item_into(index, out, labels, weights)
        return out, labels, weights, time.time() - start

    def batch(self, indices):
        """``collate_random([self[i] for i in indices])`` without the per-item tensors."""
        start = time.time()
        dev, c, B = self.device, 3 + self.channels(), len(indices)
        out = torch.empty(B, self.npoints, c, dtype=torch.float32, device=dev)
        labels = torch.empty(B, self.npoints, dtype=torch.int64, device=dev)
        weights = torch.empty(B, self.npoints, dtype=torch.float32, device=dev)
        entries = []
        for i in indices:   # the augmentation draws, item by item in the reference's order
            scene_id = self.scene_list[i]
            entries.append((self.stores[scene_id], None, self.chunk_data[scene_id], None,
                            _draw_augmentation() if self.phase == "train" else None))
        self._assemble_batch(entries, self.npoints, out, labels, weights, feat_first=True)
        return out[:, :, :3], out[:, :, 3:], labels, weights, time.time() - start


class ScannetDatasetWholeScene(_Base):
    """lib/dataset.py:252-353: tiles a scene's XY bounding box into 1.5 m columns and resamples every
    non-empty one to ``npoints`` rows.  Channel order xyz, [rgb], [normal], [multiview] (:300-311)."""

    def __init__(self, scene_list, scene_data, npoints=8192, is_weighting=True, use_color=False,
                 use_normal=False, use_multiview=False, multiview_data=None, num_classes=21, device="cuda"):
        self.npoints = npoints
        self.is_weighting = is_weighting
        self.use_multiview, self.use_color, self.use_normal = use_multiview, use_color, use_normal
        self._setup(scene_list, scene_data, multiview_data, num_classes, is_weighting, device)

    def __len__(self):
        return len(self.scene_list)

    def __getitem__(self, index):
        start = time.time()
        L, dev = self._L, self.device
        st = self.stores[self.scene_list[index]]
        cmin, cmax = st.coordmin, st.coordmax
        nx = np.ceil((cmax[0] - cmin[0]) / 1.5).astype(np.int32)
        ny = np.ceil((cmax[1] - cmin[1]) / 1.5).astype(np.int32)
        lo, hi = [], []
        for i in range(nx):
            for j in range(ny):
                curmin = cmin + [i * 1.5, j * 1.5, 0]
                curmax = cmin + [(i + 1) * 1.5, (j + 1) * 1.5, cmax[2] - cmin[2]]
                lo.append(curmin - 0.01)
                hi.append(curmax + 0.01)
        nbox = len(lo)
        lo_d = torch.from_numpy(np.stack(lo)).to(dev)
        hi_d = torch.from_numpy(np.stack(hi)).to(dev)
        sel = torch.empty(nbox, st.n, dtype=torch.int32, device=dev)
        mask = torch.empty(nbox, st.n, dtype=torch.uint8, device=dev)
        count = torch.empty(nbox, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.pn2_box_select(st.n, st.ld, st.rows.data_ptr(), nbox, lo_d.data_ptr(), hi_d.data_ptr(),
                                        sel.data_ptr(), count.data_ptr(), mask.data_ptr(), _stream(dev)),
                       "box_select")
        counts = count.tolist()                               # one device->host read per scene
        keep = [b for b in range(nbox) if counts[b] > 0]
        choices = np.stack([np.random.choice(counts[b], self.npoints, replace=True) for b in keep])
        choice = torch.from_numpy(choices.astype(np.int32)).to(dev)
        K, c = len(keep), 3 + self.channels()
        out = torch.empty(K, self.npoints, c, dtype=torch.float32, device=dev)
        labels = torch.empty(K, self.npoints, dtype=torch.int64, device=dev)
        weights = torch.empty(K, self.npoints, dtype=torch.float32, device=dev)
        self._assemble_batch([(st, sel[b], choice[t], mask[b], None) for t, b in enumerate(keep)], self.npoints,
                             out, labels, weights, feat_first=False)
        return out, labels, weights, time.time() - start


def collate_random(data):
    """lib/dataset.py:358-396 for items produced by ScannetDataset.__getitem__ (already float32 /
    int64 on the device): stack and split coords | feats."""
    point_set, semantic_seg, sample_weight, fetch_time = zip(*data)
    point_set = torch.stack(point_set)
    return (point_set[:, :, :3], point_set[:, :, 3:], torch.stack(semantic_seg), torch.stack(sample_weight),
            sum(fetch_time))
