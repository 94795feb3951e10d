"""Runs the geometric op kernels once each at the SSG level-0 shapes (for ncu captures)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import ext, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x, _, _ = synthetic.make_batch(2, B)
xyz = torch.from_numpy(np.ascontiguousarray(x[:, :, :3])).cuda()
for it in range(2):
    idx = ext.furthest_point_sampling(xyz, 1024)
    ctr = xyz.gather(1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    bq = ext.ball_query(ctr, xyz, 0.1, 32)
    d2, i3 = ext.three_nn(xyz, ctr)
torch.cuda.synchronize()
print("ok", idx.shape, bq.shape, i3.shape)
