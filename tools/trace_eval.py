"""Per-kernel totals of one eval-mode forward pass over 20 chunks (torch.profiler / CUPTI)."""
import os, sys, torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import pointnet2_semseg, synthetic
x, _, _ = synthetic.make_batch(2, int(sys.argv[1]) if len(sys.argv) > 1 else 20)
pc = torch.from_numpy(x).cuda()
torch.manual_seed(0)
model = pointnet2_semseg.get_model(20, is_msg=False, input_channels=6).cuda().eval()
with torch.no_grad():
    for _ in range(3):
        model(pc)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(pc)
        torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        d = agg.setdefault(ev.name, [0, 0.0])
        d[0] += 1
        d[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print("sum of kernel time %.1f us" % tot)
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%8.1f us %4d x  %s" % (t, c, n[:110]))
