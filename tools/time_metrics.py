"""Per-iteration metrics: the device path (pointnet2/scannet_b200/metrics.py) timed beside
the NumPy oracle of the reference's compute_acc + compute_miou on the same 32 x 8192-point batch.  (The
reference itself additionally runs a Python loop with a string hash per point, scripts/eval.py:42; measured
in the dev container: see DESIGN.md.)  Output: one JSON line."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import lib, metrics, synthetic
from oracle import metrics_np as O

NC = 20
x, y, _ = synthetic.make_batch(2, 32, 8192)
coords = np.ascontiguousarray(x[..., :3].reshape(-1, 3))
targets = (y.reshape(-1) % NC).astype(np.int64)
rng = np.random.default_rng(0)
preds = np.where(rng.random(len(targets)) < 0.5, targets, rng.integers(0, NC, len(targets))).astype(np.int64)
c, t = torch.from_numpy(coords).cuda(), torch.from_numpy(targets).cuda()
logits = torch.nn.functional.one_hot(torch.from_numpy(preds).cuda(), NC).float()
for _ in range(3):
    metrics.evaluate(c, logits, t, NC)
torch.cuda.synchronize()
lib.reset_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = metrics.evaluate(c, logits, t, NC)
e1.record()
torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 20
w = np.ones(len(targets), np.float32)
t0 = time.time()
O.compute_acc(coords, preds, targets, w, NC)
O.compute_miou(coords, preds, targets, w, NC)
cpu_ms = 1e3 * (time.time() - t0)
print(json.dumps({"workload": "Solver._eval on one 32 x 8192-point batch (arg-max, compute_acc, compute_miou)",
                  "points": int(len(targets)), "gpu_ms": round(gpu_ms, 3), "gpu_launches": lib.launch_count() // 20,
                  "host_sync": "none", "numpy_oracle_ms": round(cpu_ms, 1), "numpy_cores": 1,
                  "point_acc": float(out["point_acc"].item())}))
