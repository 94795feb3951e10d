"""FPS 8192 -> 1024 (B = 32): resident kernel vs the cloud split over a cluster of 2/4/8 CTAs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import ext, synthetic
B = 32
x, _, _ = synthetic.make_batch(2, B)
xyz = torch.from_numpy(np.ascontiguousarray(x[:, :, :3])).cuda()
xyz1k = xyz[:, :1024].contiguous()
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps, out
base, ref = timed(lambda: ext.furthest_point_sampling(xyz, 1024))
print("resident 8192->1024: %.1f us (%.0f ns/iter)" % (base, base * 1e3 / 1023))
b2, ref2 = timed(lambda: ext.furthest_point_sampling(xyz1k, 256))
print("resident 1024->256: %.1f us (%.0f ns/iter)" % (b2, b2 * 1e3 / 255))
for c in (2, 4, 8):
    os.environ["PN2_FPS_SPLIT"] = str(c)
    t, out = timed(lambda: ext.furthest_point_sampling(xyz, 1024))
    print("split %d: %.1f us (%.0f ns/iter) same=%s" % (c, t, t * 1e3 / 1023, bool((out == ref).all())))
    del os.environ["PN2_FPS_SPLIT"]
