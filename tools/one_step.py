"""Two eager SSG training steps at the benchmark size (for ncu: skip the first step's launches)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import lib, pointnet2_semseg, synthetic, train
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x, y, w = synthetic.make_batch(2, B)
batch = tuple(torch.from_numpy(a).cuda() for a in (x, y, w))
torch.manual_seed(0)
model = pointnet2_semseg.get_model(20, is_msg=False, input_channels=6).cuda().train()
step = train.TrainStep(model, lr=1e-3)
lib.reset_launch_count()
step(*batch)
torch.cuda.synchronize()
n1 = lib.launch_count()
step(*batch)
torch.cuda.synchronize()
print("library launches per step:", n1, lib.launch_count() - n1)
