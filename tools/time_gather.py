"""A/B timing (CUDA events) of the channel-major grouping / interpolation forward kernels at the SSG and
multiview shapes: the L1-gather kernels (PN2_GROUP_SMEM=0 / PN2_INTERP_SMEM=0) against the shared-memory
gather that is the default, with the algorithmic bytes of SURVEY.md 8(d) and the measured HBM peak.

    python tools/time_gather.py [B]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import ext  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REPS = 20
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    PEAK = 6650.0
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2 (126 MB)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(REPS):
        FLUSH.zero_()  # cold L2 for every repetition
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return 1e3 * total / REPS  # us


def ab(name, env, nbytes, fn):
    os.environ[env] = "0"          # the L1-gather kernel
    base = fn()
    t0 = timed(fn)
    os.environ.pop(env, None)      # default: the shared-memory gather where the shape suits it
    same = torch.equal(fn(), base)
    t1 = timed(fn)
    print("%-44s L1 gather (%s=0) %8.1f us %5.1f%%   shared-memory gather (default) %8.1f us %5.1f%%   bit-equal %s" % (
        name, env, t0, 100 * nbytes / t0 / 1e3 / PEAK, t1, 100 * nbytes / t1 / 1e3 / PEAK, same))


rng = np.random.default_rng(0)
# grouping: (C, N, npoint, nsample) of the SSG SA levels' feature grouping + multiview SA1
for C, N, npoint, ns in [(6, 8192, 1024, 32), (64, 1024, 256, 32), (128, 256, 64, 32), (256, 64, 16, 32),
                         (131, 8192, 1024, 32)]:
    feats = torch.randn(B, C, N, device="cuda")
    idx = torch.from_numpy(rng.integers(0, N, (B, npoint, ns)).astype(np.int32)).cuda()
    E = npoint * ns
    nbytes = 4 * B * E + 4 * B * C * min(N, E) + 4 * B * C * E
    ab("group C=%d N=%d -> %dx%d" % (C, N, npoint, ns), "PN2_GROUP_SMEM", nbytes,
       lambda: ext.group_points(feats, idx))
# interpolation: (C, m, n) of the SSG FP levels
for C, m, n in [(128, 1024, 8192), (256, 256, 1024), (256, 64, 256), (512, 16, 64)]:
    feats = torch.randn(B, C, m, device="cuda")
    idx = torch.from_numpy(rng.integers(0, m, (B, n, 3)).astype(np.int32)).cuda()
    w = torch.rand(B, n, 3, device="cuda")
    w = (w / w.sum(dim=2, keepdim=True)).contiguous()
    nbytes = 24 * B * n + 4 * B * C * min(m, 3 * n) + 4 * B * C * n
    ab("interpolate C=%d m=%d -> n=%d" % (C, m, n), "PN2_INTERP_SMEM", nbytes,
       lambda: ext.three_interpolate(feats, idx, w))
