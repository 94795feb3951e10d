"""Runs each rows-layout kernel once at an SSG-sized shape (for ncu captures): a fused SA block
(B x 256 centres x 32 samples, 64->64->128) and an FP block, forward + backward."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import pointnet2_modules as M, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
x, _, _ = synthetic.make_batch(2, B, npoints=1024)
xyz = torch.from_numpy(np.ascontiguousarray(x[:, :, :3])).cuda()
feats = torch.randn(B, 64, 1024, device="cuda", requires_grad=True)
sa = M.PointnetSAModule(npoint=256, radius=0.2, nsample=32, mlp=[64, 64, 64, 128]).cuda().train()
fp = M.PointnetFPModule(mlp=[128 + 64, 128, 128]).cuda().train()
for it in range(2):
    new_xyz, f = sa(xyz, feats)
    up = fp(xyz, new_xyz, feats, f)
    up.square().mean().backward()
torch.cuda.synchronize()
print("ok", f.shape, up.shape)
