"""Chunk generation: the CUDA path (pointnet2/scannet_b200/chunker.py) timed beside the
reference's NumPy algorithm (oracle/chunker_np.py, pinned against lib/dataset.py) on the same synthetic
scenes: 32 scenes x 150k points -> one epoch's generate_chunks() + one training batch of 32 x 8192-pt
chunks (xyz + rgb + normal, augmented), and one whole-scene tiling.  Output: one JSON line."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import chunker, lib, synthetic
from oracle import chunker_np as O

NS, NP, NPTS = int(os.environ.get("SCENES", 32)), 150000, 8192
scenes = {"s%02d" % i: synthetic.make_prep_scene(100 + i, npoints=NP, size=(6.0, 5.0, 2.7)) for i in range(NS)}
names = sorted(scenes)
flags = dict(use_color=True, use_normal=True)
t0 = time.time()
ds = chunker.ScannetDataset("train", names, scenes, npoints=NPTS, **flags)
torch.cuda.synchronize()
upload_s = time.time() - t0


def gpu_epoch():
    ds.generate_chunks()
    out = ds.batch(range(NS))
    torch.cuda.synchronize()
    return out


np.random.seed(0)
gpu_epoch()
lib.reset_launch_count()
times = []
for r in range(5):
    np.random.seed(r)
    t0 = time.time(); gpu_epoch(); times.append(time.time() - t0)
gpu_s = float(np.median(times))
launches = lib.launch_count() // 5
t0 = time.time(); b = ds.batch(range(NS)); torch.cuda.synchronize(); batch_s = time.time() - t0

lw = O.label_weights([scenes[k] for k in names], 21)
np.random.seed(0)
t0 = time.time()
chunks = [O.generate_chunk(scenes[k], NPTS, np.random)[0] for k in names]
items = [O.get_item(c, lw, True, rng=np.random, **flags)[:3] for c in chunks]
O.collate(items)
cpu_s = time.time() - t0

ws = chunker.ScannetDatasetWholeScene(names[:1], scenes, npoints=NPTS, **flags)
np.random.seed(1); ws[0]; torch.cuda.synchronize()
t0 = time.time(); np.random.seed(1); pts = ws[0][0]; torch.cuda.synchronize(); ws_gpu = time.time() - t0
t0 = time.time(); np.random.seed(1); O.whole_scene_tiles(scenes[names[0]], lw, NPTS, rng=np.random, **flags); ws_cpu = time.time() - t0
print(json.dumps({
    "workload": "%d scenes x %d pts -> generate_chunks + one batch of %d x %d-pt chunks (xyz+rgb+normal, augmented)" % (NS, NP, NS, NPTS),
    "gpu_chunks_per_s": round(NS / gpu_s, 1), "gpu_ms": round(1e3 * gpu_s, 2), "gpu_batch_only_ms": round(1e3 * batch_s, 2),
    "gpu_launches": launches, "scene_upload_s": round(upload_s, 2),
    "numpy_chunks_per_s": round(NS / cpu_s, 1), "numpy_ms": round(1e3 * cpu_s, 1), "numpy_cores": 1,
    "whole_scene_tiles": int(pts.shape[0]), "whole_scene_gpu_ms": round(1e3 * ws_gpu, 2), "whole_scene_numpy_ms": round(1e3 * ws_cpu, 1),
}))
