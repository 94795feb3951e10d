"""CUDA-event timing of the nine op kernels at the SSG level shapes (B x 8192)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import ext, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REPS = 10


def timed(name, fn):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-52s %9.1f us" % (name, 1e3 * e0.elapsed_time(e1) / REPS))
    return out


x, _, _ = synthetic.make_batch(2, B)
xyz = torch.from_numpy(np.ascontiguousarray(x[:, :, :3])).cuda()
levels = [(8192, 1024, 0.1), (1024, 256, 0.2), (256, 64, 0.4), (64, 16, 0.8)]
cur = xyz
pyr = [xyz]
for n, m, r in levels:
    idx = timed("fps %d->%d" % (n, m), lambda: ext.furthest_point_sampling(cur, m))
    ctr = cur.gather(1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    timed("ball_query n=%d m=%d r=%.1f ns=32" % (n, m, r), lambda: ext.ball_query(ctr, cur, r, 32))
    timed("three_nn n=%d m=%d" % (n, m), lambda: ext.three_nn(cur, ctr))
    cur = ctr
    pyr.append(ctr)
